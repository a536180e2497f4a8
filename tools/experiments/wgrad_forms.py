#!/usr/bin/env python
"""Which formulation of the weight-gradient GEMM (K = rows = 24 100 .. 153 000, M, N in {256, 2048}) does cuBLAS run fastest?
g.t() @ x (what autograd's Linear does) picks an sm_80 split-K kernel on B200 (profiles/r1_encoder_kernel_breakdown.txt)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
dev = torch.device("cuda:0")
torch.backends.cuda.matmul.allow_tf32 = True


def timeit(f, iters=50):
    for _ in range(5):
        f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        f()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3


for rows in (24100, 153000):
    for cout, cin in ((256, 256), (2048, 256), (256, 2048), (384, 256)):
        g, x = torch.randn(rows, cout, device=dev), torch.randn(rows, cin, device=dev)
        gt, xt = g.t().contiguous(), x.t().contiguous()
        out = torch.empty(cout, cin, device=dev)
        forms = {
            "g.t() @ x": lambda: g.t() @ x,
            "(x.t() @ g).t()": lambda: (x.t() @ g).t(),
            "mm(out=)": lambda: torch.mm(g.t(), x, out=out),
            "einsum": lambda: torch.einsum("ro,ri->oi", g, x),
            "contig operands (copies not timed)": lambda: gt @ x,
            "both transposed contig": lambda: gt @ xt.t(),
        }
        ref = (g.double().t() @ x.double())
        res = {}
        for k, f in forms.items():
            y = f()
            err = float((y.double() - ref).abs().max() / ref.abs().max())
            res[k] = (round(timeit(f), 1), f"{err:.1e}")
        print(json.dumps({"rows": rows, "out": cout, "in": cin, "us (tf32), rel err": res}), flush=True)
