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


ONLY = set(sys.argv[1:])            # e.g. `bench_epilogue.py layout` runs the flatten / decoder-consumer rows alone
for rows in (() if ONLY and "epilogue" not in ONLY else (24100, 153000)):
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

# ---- dropout variants (training) ----
for rows in (() if ONLY and "dropout" not in ONLY else (24100, 153000)):
    C = 256
    nset = max(2, int(400e6 // (rows * C * 4 * 4)) + 1)
    sets = [dict(x=torch.randn(rows, C, device=dev), r=torch.randn(rows, C, device=dev), dy=torch.randn(rows, C, device=dev)) for _ in range(nset)]
    gamma, beta, bias = torch.ones(C, device=dev), torch.zeros(C, device=dev), torch.randn(C, device=dev)
    z, y = torch.empty(rows, C, device=dev), torch.empty(rows, C, device=dev)
    mean, rstd = torch.empty(rows, device=dev), torch.empty(rows, device=dev)
    dz, dx = torch.empty(rows, C, device=dev), torch.empty(rows, C, device=dev)
    dg, db, dbi = torch.empty(C, device=dev), torch.empty(C, device=dev), torch.empty(C, device=dev)
    rng = epilogue.new_rng(dev)
    st = lambda: torch.cuda.current_stream().cuda_stream
    fwd = [lambda s=s: L.msda_epilogue_ln_dropout_forward_f32(s["x"].data_ptr(), bias.data_ptr(), s["r"].data_ptr(), gamma.data_ptr(), beta.data_ptr(), 1e-5, rows, C, rng.data_ptr(), 1, 0.1, z.data_ptr(), y.data_ptr(), mean.data_ptr(), rstd.data_ptr(), st()) for s in sets]
    t_fwd = [lambda s=s: F.layer_norm(s["r"] + F.dropout(s["x"] + bias, 0.1, True), (C,), gamma, beta) for s in sets]
    report("epilogue_ln_dropout_fwd", rows, C, rows * C * 4 * 4, timeit(fwd), timeit(t_fwd))
    fwd[0]()
    bwd = [lambda s=s: L.msda_epilogue_ln_dropout_backward_f32(s["dy"].data_ptr(), s["x"].data_ptr(), mean.data_ptr(), rstd.data_ptr(), gamma.data_ptr(), rows, C, rng.data_ptr(), 1, 0.1, dz.data_ptr(), dx.data_ptr(), dg.data_ptr(), db.data_ptr(), dbi.data_ptr(), st()) for s in sets]
    keep = [torch.rand(rows, C, device=dev) > 0.1 for _ in sets[:2]]

    def torch_bwd_drop(s, k):
        g = torch.ops.aten.native_layer_norm_backward(s["dy"], s["x"], [C], mean.view(-1, 1), rstd.view(-1, 1), gamma, beta, [True, True, True])
        d = torch.ops.aten.native_dropout_backward(g[0], k, 1.0 / 0.9)
        return g, d, d.sum(0)
    report("epilogue_ln_dropout_bwd", rows, C, rows * C * 4 * 4, timeit(bwd), timeit([lambda s=s, k=keep[i % 2]: torch_bwd_drop(s, k) for i, s in enumerate(sets)]))
    del sets, keep
    C = 2048
    hs = [torch.relu(torch.randn(rows, C, device=dev)) for _ in range(2 if rows > 100000 else 3)]
    dr = [lambda h=h: L.msda_dropout_inplace_f32(h.data_ptr(), h.numel(), rng.data_ptr(), 2, 0.1, st()) for h in hs]
    report("dropout_inplace", rows, C, rows * C * 4 * 2, timeit(dr, 20), timeit([lambda h=h: F.dropout(h, 0.1, True) for h in hs], 20))
    del hs

# ---- caller-side flattening and the decoder-side consumers (rank 3) ----
from ocpg_b200 import decoder as dec_mod, flatten as flat_mod
from ocpg_b200.workloads import A2D_ENCODER, YTVOS_ENCODER
for wl in (A2D_ENCODER, YTVOS_ENCODER):
    N, C, S = wl.n_frames, 256, wl.S
    nset = max(2, int(300e6 // (N * S * C * 4 * 4)) + 1)
    sets = [dict(src=[torch.randn(N, C, h, w, device=dev) for h, w in wl.levels], pos=[torch.randn(N, C, h, w, device=dev) for h, w in wl.levels],
                 mem=torch.randn(N, S, C, device=dev)) for _ in range(nset)]
    le = torch.randn(len(wl.levels), C, device=dev)

    def torch_flatten(s):
        a = torch.cat([t.flatten(2).transpose(1, 2) for t in s["src"]], 1)
        b = torch.cat([p.flatten(2).transpose(1, 2) + le[l].view(1, 1, -1) for l, p in enumerate(s["pos"])], 1)
        return a, b

    def torch_unflatten(s):
        out, at = [], 0
        for h, w in wl.levels[:-1]:
            out.append(s["mem"][:, at:at + h * w, :].reshape(N, h, w, C).permute(0, 3, 1, 2).contiguous())
            at += h * w
        return out
    report(f"flatten_levels[{wl.name}]", N * S, C, N * S * C * 4 * 4, timeit([lambda s=s: flat_mod._flatten_native(s["src"], s["pos"], le) for s in sets]),
           timeit([lambda s=s: torch_flatten(s) for s in sets]))
    cov = sum(h * w for h, w in wl.levels[:-1])
    report(f"unflatten_levels[{wl.name}]", N * cov, C, N * cov * C * 4 * 2, timeit([lambda s=s: flat_mod._unflatten_native(s["mem"], wl.levels[:-1]) for s in sets]),
           timeit([lambda s=s: torch_unflatten(s) for s in sets]))
    del sets
N, Lq, M, Lv, P = 5, 5, 8, 4, 4
loc = torch.rand(N, Lq, M, Lv, P, 2, device=dev)
aw = torch.softmax(torch.randn(N, Lq, M * Lv * P, device=dev), -1).view(N, Lq, M, Lv, P)
vr = 0.8 + 0.2 * torch.rand(N, Lv, 2, device=dev)
ref = torch.rand(N, Lq, 2, device=dev)


def torch_consumers():
    rpi = ref[:, :, None] * vr[:, None]
    sl = loc / vr[:, None, None, :, None, :]
    tw, ti = aw.view(N, Lq, -1).topk(30, dim=2)
    return rpi, torch.gather(sl.view(N, Lq, -1, 2), 2, ti.unsqueeze(-1).repeat(1, 1, 1, 2))


us = timeit([lambda: (dec_mod.scale_reference_points(ref, vr), dec_mod.select_top_samples(loc, aw, vr, 30))])
us_t = timeit([torch_consumers])
print(json.dumps(dict(kernel="decoder consumers (reference points + top-30 samples), config 4: 25 queries", us=round(us, 2),
                      torch_us=round(us_t, 2), speedup_vs_torch=round(us_t / us, 2), note="device time of the launches, graph replay")), flush=True)
