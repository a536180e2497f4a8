#!/usr/bin/env python
"""HBM roofline of the encoder-layer epilogue kernels (ocpg_b200/epilogue.py): algorithmic bytes / device time vs the
measured copy bandwidth (MEASURED_PEAKS.json), with the torch operators they replace timed beside them.  Inputs rotate
through enough buffers to exceed L2 between reuses.  One JSON line per kernel and shape."""
import json, os, statistics, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.nn.functional as F
import ocpg_b200
from ocpg_b200 import epilogue

dev = torch.device("cuda:0")
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    PEAK = 6650.0
L = ocpg_b200.lib()


def timeit(fns, iters=40):
    """Mean device time per call: every call captured in its own CUDA graph (no host launch latency between kernels)."""
    with torch.cuda.stream(torch.cuda.Stream()):
        for f in fns:
            f()
    torch.cuda.synchronize()
    graphs, keep = [], []
    for f in fns:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            keep.append(f())
        graphs.append(g)
    for g in graphs:
        g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(iters):
        graphs[i % len(graphs)].replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3      # us


def report(name, rows, C, nbytes, us, us_torch):
    rec = dict(kernel=name, rows=rows, channels=C, algorithmic_MB=round(nbytes / 1e6, 1), us=round(us, 2),
               gbs=round(nbytes / us / 1e3, 1), frac_of_hbm_peak=round(nbytes / us / 1e3 / PEAK, 3), torch_us=round(us_torch, 2),
               speedup_vs_torch=round(us_torch / us, 2))
    print(json.dumps(rec), flush=True)


for rows in (24100, 153000):
    C = 256
    nset = max(2, int(400e6 // (rows * C * 4 * 4)) + 1)
    sets = [dict(x=torch.randn(rows, C, device=dev), r=torch.randn(rows, C, device=dev), dy=torch.randn(rows, C, device=dev)) for _ in range(nset)]
    gamma, beta, bias = torch.ones(C, device=dev), torch.zeros(C, device=dev), torch.randn(C, device=dev)
    z, y = torch.empty(rows, C, device=dev), torch.empty(rows, C, device=dev)
    mean, rstd = torch.empty(rows, device=dev), torch.empty(rows, device=dev)
    dz, dg, db, dbi = torch.empty(rows, C, device=dev), torch.empty(C, device=dev), torch.empty(C, device=dev), torch.empty(C, device=dev)
    fwd = [lambda s=s: L.msda_epilogue_ln_forward_f32(s["x"].data_ptr(), bias.data_ptr(), s["r"].data_ptr(), gamma.data_ptr(), beta.data_ptr(), 1e-5, rows, C, z.data_ptr(), y.data_ptr(), mean.data_ptr(), rstd.data_ptr(), torch.cuda.current_stream().cuda_stream) for s in sets]
    t_fwd = [lambda s=s: F.layer_norm(s["r"] + (s["x"] + bias), (C,), gamma, beta) for s in sets]
    report("epilogue_ln_fwd", rows, C, rows * C * 4 * 4, timeit(fwd), timeit(t_fwd))
    fwd[0]()
    bwd = [lambda s=s: L.msda_epilogue_ln_backward_f32(s["dy"].data_ptr(), s["x"].data_ptr(), mean.data_ptr(), rstd.data_ptr(), gamma.data_ptr(), rows, C, dz.data_ptr(), dg.data_ptr(), db.data_ptr(), dbi.data_ptr(), torch.cuda.current_stream().cuda_stream) for s in sets]

    def torch_bwd(s):
        g = torch.ops.aten.native_layer_norm_backward(s["dy"], s["x"], [C], mean.view(-1, 1), rstd.view(-1, 1), gamma, beta, [True, True, True])
        return g, g[0].sum(0)
    report("epilogue_ln_bwd", rows, C, rows * C * 4 * 3, timeit(bwd), timeit([lambda s=s: torch_bwd(s) for s in sets]))
    cs = [lambda s=s: epilogue.column_sum(s["x"]) for s in sets]
    report("column_sum", rows, C, rows * C * 4, timeit(cs), timeit([lambda s=s: s["x"].sum(0) for s in sets]))
    del sets
    C = 2048
    nset = 2 if rows > 100000 else 3
    sets = [dict(dh=torch.randn(rows, C, device=dev), h=torch.relu(torch.randn(rows, C, device=dev))) for _ in range(nset)]
    dpre, dbias = torch.empty(rows, C, device=dev), torch.empty(C, device=dev)
    rb = [lambda s=s: L.msda_relu_backward_column_sum_f32(s["dh"].data_ptr(), s["h"].data_ptr(), rows, C, dpre.data_ptr(), dbias.data_ptr(), torch.cuda.current_stream().cuda_stream) for s in sets]

    def torch_rb(s):
        d = torch.ops.aten.threshold_backward(s["dh"], s["h"], 0)
        return d, d.sum(0)
    report("relu_backward_column_sum", rows, C, rows * C * 4 * 3, timeit(rb, 20), timeit([lambda s=s: torch_rb(s) for s in sets], 20))
    del sets, dpre
