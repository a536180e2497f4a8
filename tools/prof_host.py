#!/usr/bin/env python
"""cProfile of the eager encoder step's HOST side (A2D shape, 5 frames): where the Python time of the fused path goes."""
import cProfile, io, os, pstats, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ocpg_b200.encoder import build_encoder
from ocpg_b200.workloads import A2D_ENCODER as wl
dev = torch.device("cuda:0")
torch.backends.cuda.matmul.allow_tf32 = True
enc = build_encoder(dropout=float(sys.argv[1]) if len(sys.argv) > 1 else 0.0).to(dev).train()
shapes = torch.tensor(wl.levels, dtype=torch.int64, device=dev)
start = torch.cat((shapes.new_zeros(1), shapes.prod(1).cumsum(0)[:-1]))
src = torch.randn(5, wl.S, 256, device=dev); pos = torch.randn(5, wl.S, 256, device=dev) * 0.1
vr = torch.ones(5, 4, 2, device=dev); g = torch.randn(5, wl.S, 256, device=dev)


def step():
    for p in enc.parameters():
        p.grad = None
    x = src.clone().requires_grad_(True)
    enc(x, shapes, start, vr, pos, None).backward(g)


for _ in range(5):
    step()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(20):
    step()
pr.disable()
torch.cuda.synchronize()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(28)
print(s.getvalue())
