#!/usr/bin/env python
"""Pinned host <-> device copy bandwidth on this box: one direction, both at once, and in 1 vs 4 pieces per step -- the
ceiling of bench.py's e2e leg (86.4 MB in + 86.4 MB out per step)."""
import json, time, torch
dev = torch.device("cuda:0")
n = 86_374_400 // 4
h_in, h_out = torch.empty(n, dtype=torch.float32).pin_memory(), torch.empty(n, dtype=torch.float32).pin_memory()
d_in, d_out = torch.empty(n, device=dev), torch.empty(n, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h, pieces, iters=30):
    cuts = [(i * n // pieces, (i + 1) * n // pieces) for i in range(pieces)]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(iters):
        if h2d:
            with torch.cuda.stream(s1):
                for a, b in cuts:
                    d_in[a:b].copy_(h_in[a:b], non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                for a, b in cuts:
                    h_out[a:b].copy_(d_out[a:b], non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / iters
    return round(n * 4 / dt / 1e9, 1), round(dt * 1e3, 3)


for name, a, b, p in (("h2d only", 1, 0, 1), ("d2h only", 0, 1, 1), ("both", 1, 1, 1), ("both, 4 pieces", 1, 1, 4), ("both, 16 pieces", 1, 1, 16)):
    run(a, b, p, 3)
    gbs, ms = run(a, b, p)
    print(json.dumps({"case": name, "GB/s per direction": gbs, "ms per 86.4 MB step": ms}), flush=True)
