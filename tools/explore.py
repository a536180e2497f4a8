#!/usr/bin/env python
"""Kernel exploration on the GPU box: times our forward / backward under the library's tuning knobs,
the generic kernels and the reference's own CUDA op (oracle/_ref) on the BASELINE.json shapes, in both
location regimes.  Prints one JSON object per measurement (and a table) -- development tool, not the
benchmark of record (that is bench.py)."""
import argparse
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import ocpg_b200  # noqa: E402
import ocpg_b200.MultiScaleDeformableAttention as MSDA  # noqa: E402
from ocpg_b200.workloads import A2D_DECODER, A2D_ENCODER, YTVOS_ENCODER, make_inputs  # noqa: E402


def timeit(fn, sets, iters, flush=None):
    for i in range(3):
        fn(sets[i % len(sets)])
    torch.cuda.synchronize()
    ts = []
    for i in range(iters):
        if flush is not None:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(sets[i % len(sets)]); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return statistics.median(ts), min(ts)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=30)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "explore.jsonl"))
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    fout = open(args.out, "a")
    ref = None
    try:
        from oracle import build_ref_cuda
        ref = build_ref_cuda.load()
    except Exception as e:
        print("reference CUDA op unavailable:", e)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > L2 (126 MB)
    rows = []

    def emit(**kw):
        rows.append(kw)
        fout.write(json.dumps(kw) + "\n"); fout.flush()
        print(json.dumps(kw), flush=True)

    workloads = [A2D_ENCODER, YTVOS_ENCODER, A2D_DECODER] if not args.quick else [A2D_ENCODER]
    for wl in workloads:
        fb, bb = wl.algorithmic_bytes()
        for regime in ("init", "uniform"):
            nsets = 3 if wl is YTVOS_ENCODER else 4
            sets = [make_inputs(wl, regime, seed=i, device=dev) for i in range(nsets)]
            fwd = lambda x: MSDA.ms_deform_attn_forward(x["value"], x["shapes"], x["start"], x["loc"], x["attn"], 64)
            bwd = lambda x: MSDA.ms_deform_attn_backward(x["value"], x["shapes"], x["start"], x["loc"], x["attn"], x["grad_out"], 64)

            def measure(tag, **opts):
                for k, v in opts.items():
                    ocpg_b200.set_option(k, v)
                try:
                    for name, fn, nbytes in (("fwd", fwd, fb), ("bwd", bwd, bb)):
                        for cold in (True, False):
                            med, mn = timeit(fn, sets, args.iters, flush if cold else None)
                            emit(workload=wl.name, regime=regime, impl="ours", variant=tag, op=name, l2="cold" if cold else "warm",
                                 us_median=round(med, 2), us_min=round(mn, 2), gbs=round(nbytes / med / 1e3, 1),
                                 frac_of_6551=round(nbytes / med / 1e3 / 6551, 4))
                finally:
                    for k in opts:
                        ocpg_b200.set_option(k, 0)

            measure("default")
            if wl is not A2D_DECODER:
                measure("linear_walk", force_linear_walk=1)
                measure("bwd_query_major", bwd_algo=1)
                measure("fwd1cta_bwd1cta", fwd_ctas_per_sm=1, bwd_ctas_per_sm=1)
                if not args.quick:
                    measure("fwd3", fwd_ctas_per_sm=3)
                    measure("bwd2", bwd_ctas_per_sm=2)
            if wl is A2D_ENCODER:
                measure("generic_kernels", force_generic=1)
            # bf16 value
            bsets = [dict(x, value=x["value"].bfloat16(), grad_out=x["grad_out"].bfloat16()) for x in sets[:2]]
            fb16, bb16 = wl.algorithmic_bytes(2, 2)
            for name, fn, nbytes in (("fwd", fwd, fb16), ("bwd", bwd, bb16)):
                med, mn = timeit(fn, bsets, args.iters, flush)
                emit(workload=wl.name, regime=regime, impl="ours", variant="bf16", op=name, l2="cold", us_median=round(med, 2),
                     us_min=round(mn, 2), gbs=round(nbytes / med / 1e3, 1), frac_of_6551=round(nbytes / med / 1e3 / 6551, 4))
            del bsets
            if ref is not None and wl.n_frames <= 64:
                rf = lambda x: ref.ms_deform_attn_forward(x["value"], x["shapes"], x["start"], x["loc"], x["attn"], 64)
                rb = lambda x: ref.ms_deform_attn_backward(x["value"], x["shapes"], x["start"], x["loc"], x["attn"], x["grad_out"], 64)
                for name, fn, nbytes in (("fwd", rf, fb), ("bwd", rb, bb)):
                    for cold in (True, False):
                        med, mn = timeit(fn, sets, max(5, args.iters // 3), flush if cold else None)
                        emit(workload=wl.name, regime=regime, impl="reference_cuda_sm100a", variant="stock", op=name,
                             l2="cold" if cold else "warm", us_median=round(med, 2), us_min=round(mn, 2),
                             gbs=round(nbytes / med / 1e3, 1), frac_of_6551=round(nbytes / med / 1e3 / 6551, 4))
            del sets
            torch.cuda.empty_cache()
    print("\n%-34s %-8s %-22s %-26s %-4s %-5s %10s %10s %8s" % ("workload", "regime", "impl", "variant", "op", "l2", "us_med", "us_min", "frac"))
    for r in rows:
        print("%-34s %-8s %-22s %-26s %-4s %-5s %10.1f %10.1f %8.3f" % (r["workload"], r["regime"], r["impl"], r["variant"],
              r["op"], r["l2"], r["us_median"], r["us_min"], r["frac_of_6551"]))


if __name__ == "__main__":
    main()
