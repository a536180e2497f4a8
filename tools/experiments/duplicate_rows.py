#!/usr/bin/env python
"""How many of the backward's grad_value reds hit a row that another red of the same query / warp (4 x-adjacent queries) /
CTA tile (8 x 4 queries) also hits, per head and pyramid level, in the init regime (sigma = 2 px)?  CPU only."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, numpy as np
from ocpg_b200.workloads import A2D_ENCODER as wl, make_inputs
x = make_inputs(wl, "init", seed=0)
loc = x["loc"][0]            # (Lq, M, L, P, 2)
Lq, M, L, P, _ = loc.shape
shapes = x["shapes"].tolist(); start = x["start"].tolist()
rows = torch.full((Lq, M, L, P, 4), -1, dtype=torch.int64)
for l, (H, W) in enumerate(shapes):
    px = loc[:, :, l, :, 0] * W - 0.5; py = loc[:, :, l, :, 1] * H - 0.5
    inr = (py > -1) & (px > -1) & (py < H) & (px < W)
    x0 = torch.floor(px).long(); y0 = torch.floor(py).long()
    for c, (dy, dx) in enumerate(((0, 0), (0, 1), (1, 0), (1, 1))):
        xx, yy = x0 + dx, y0 + dy
        ok = inr & (xx >= 0) & (xx < W) & (yy >= 0) & (yy < H)
        r = start[l] + yy * W + xx
        rows[:, :, l, :, c] = torch.where(ok, r, torch.full_like(r, -1))
# level-0 queries only, groups of 4 x-adjacent (W0 = 80 divisible by 4)
H0, W0 = shapes[0]
q0 = rows[:H0 * W0].view(H0, W0 // 4, 4, M, L, P * 4)          # (y, xg, 4 queries, M, L, 16 items)
valid = (q0 >= 0)
tot_valid = valid.sum().item()
# type (i): distinct within (query, head, level)
def ndistinct(t):       # t: (..., K) with -1 invalid -> count distinct valid per leading index, summed
    s, _ = torch.sort(t, dim=-1)
    d = (s[..., 1:] != s[..., :-1]) & (s[..., 1:] >= 0)
    first = (s[..., :1] >= 0)
    return (d.sum(-1) + first.squeeze(-1).long()).sum().item()
per_query = ndistinct(q0)
per_warp = ndistinct(q0.permute(0, 1, 3, 4, 2, 5).reshape(H0, W0 // 4, M, L, 64))
print("valid items", tot_valid, "distinct per (query,head,level)", per_query, "-> dup frac", 1 - per_query / tot_valid)
print("distinct per (warp = 4 queries, head, level)", per_warp, "-> dup frac", 1 - per_warp / tot_valid)
for l in range(L):
    v = (q0[..., l, :] >= 0).sum().item()
    a = ndistinct(q0[..., l, :]); b = ndistinct(q0[..., l, :].permute(0, 1, 3, 2, 4).reshape(H0, W0 // 4, M, 64))
    print(f"level {l}: valid {v}  per-query dup {1 - a / v:.3f}  per-warp dup {1 - b / v:.3f}")
# CTA tile 8x4 queries (8 warps): distinct per (tile, head, level)
t = rows[:H0 * W0].view(H0 // 1, W0, M, L, 16)[: (H0 // 4) * 4].view(H0 // 4, 4, W0 // 8, 8, M, L, 16).permute(0, 2, 4, 5, 1, 3, 6).reshape(H0 // 4, W0 // 8, M, L, 512)
vt = (t >= 0).sum().item()
print("per CTA tile (32 queries) dup frac", 1 - ndistinct(t) / vt)
