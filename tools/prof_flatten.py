#!/usr/bin/env python
"""One flatten + one unflatten launch on the YTVOS shape, for `ncu --set full` (tools/prof_flatten.py)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ocpg_b200 import flatten as flat_mod
from ocpg_b200.workloads import YTVOS_ENCODER as wl
dev = torch.device("cuda:0")
N, C = wl.n_frames, 256
src = [torch.randn(N, C, h, w, device=dev) for h, w in wl.levels]
pos = [torch.randn(N, C, h, w, device=dev) for h, w in wl.levels]
le = torch.randn(len(wl.levels), C, device=dev)
for _ in range(2):
    s, p = flat_mod._flatten_native(src, pos, le)
    maps = flat_mod._unflatten_native(s, wl.levels[:-1])
torch.cuda.synchronize()
