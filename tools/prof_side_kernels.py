#!/usr/bin/env python
"""One launch of each kernel around the operator (epilogue incl. dropout, flatten / unflatten, decoder consumers) on the
YTVOS / decoder shapes, for ncu (profiles/r1_ncu_side_kernels.txt)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from ocpg_b200 import decoder as dec_mod, epilogue, flatten as flat_mod
from ocpg_b200.workloads import YTVOS_ENCODER as wl
dev = torch.device("cuda:0")
N, C, S = wl.n_frames, 256, wl.S
src = [torch.randn(N, C, h, w, device=dev) for h, w in wl.levels]
pos = [torch.randn(N, C, h, w, device=dev) for h, w in wl.levels]
le = torch.randn(len(wl.levels), C, device=dev)
x = torch.randn(N, S, C, device=dev, requires_grad=True)
res = torch.randn(N, S, C, device=dev, requires_grad=True)
w1, b1 = (torch.randn(2048, C, device=dev) * 0.05).requires_grad_(True), torch.randn(2048, device=dev, requires_grad=True)
gamma, beta, bias = (torch.randn(C, device=dev, requires_grad=True) for _ in range(3))
for _ in range(2):
    s, p = flat_mod._flatten_native(src, pos, le)
    flat_mod._unflatten_native(s, wl.levels[:-1])
    rng = epilogue.new_rng(dev)
    for pdrop in (0.0, 0.1):
        y = epilogue.bias_residual_layer_norm(x, bias, res, gamma, beta, 1e-5, rng if pdrop else None, 1, pdrop)
        h = epilogue.linear_relu(y, w1, b1, rng if pdrop else None, 2, pdrop)
        (y.sum() + h.sum()).backward()
    loc = torch.rand(5, 5, 8, 4, 4, 2, device=dev)
    aw = torch.softmax(torch.randn(5, 5, 128, device=dev), -1).view(5, 5, 8, 4, 4)
    vr = torch.ones(5, 4, 2, device=dev)
    dec_mod.select_top_samples(loc, aw, vr, 30)
    dec_mod.scale_reference_points(torch.rand(5, 5, 2, device=dev), vr)
torch.cuda.synchronize()
