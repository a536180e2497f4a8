#!/usr/bin/env python
"""Time forward / backward under lists of library options (development tool; bench.py is the benchmark of record).

    python tools/sweep.py --workload a2d --regime init --variants "w8:fwd_warps=8,bwd_warps=8;w16:fwd_warps=16,bwd_warps=16"
"""
import argparse, json, os, statistics, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import ocpg_b200  # noqa: E402
import ocpg_b200.MultiScaleDeformableAttention as MSDA  # noqa: E402
from ocpg_b200.workloads import A2D_DECODER, A2D_ENCODER, YTVOS_ENCODER, make_inputs  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="a2d")
ap.add_argument("--regime", default="init")
ap.add_argument("--dtype", default="f32")
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--variants", default="default:")
ap.add_argument("--ops", default="fwd,bwd")
ap.add_argument("--sigma", type=float, default=2.0, help="init regime: std of the sampling offsets in pixels")
ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "sweep.jsonl"))
args = ap.parse_args()
wl = {"a2d": A2D_ENCODER, "ytvos": YTVOS_ENCODER, "decoder": A2D_DECODER}[args.workload]
dev = torch.device("cuda:0")
vdt = torch.bfloat16 if args.dtype == "bf16" else None
nsets = 2 if wl is YTVOS_ENCODER else 4
sets = [make_inputs(wl, args.regime, seed=i, device=dev, value_dtype=vdt, sigma_px=args.sigma) for i in range(nsets)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
vb = 2 if vdt is not None else 4
fb, bb = wl.algorithmic_bytes(vb, vb)
fns = {"fwd": (lambda x: MSDA.ms_deform_attn_forward(x["value"], x["shapes"], x["start"], x["loc"], x["attn"], 64), fb),
       "bwd": (lambda x: MSDA.ms_deform_attn_backward(x["value"], x["shapes"], x["start"], x["loc"], x["attn"], x["grad_out"], 64), bb)}
fout = open(args.out, "a")
for var in args.variants.split(";"):
    name, _, optstr = var.partition(":")
    opts = dict((kv.split("=")[0], int(kv.split("=")[1])) for kv in optstr.split(",") if kv)
    for k, v in opts.items():
        ocpg_b200.set_option(k, v)
    for op in args.ops.split(","):
        fn, nbytes = fns[op]
        for i in range(3):
            fn(sets[i % nsets])
        torch.cuda.synchronize()
        ts = []
        for i in range(args.iters):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(sets[i % nsets]); b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) * 1e3)
        med = statistics.median(ts)
        rec = dict(workload=wl.name, regime=args.regime, sigma_px=args.sigma, dtype=args.dtype, variant=name, opts=opts, op=op, us_median=round(med, 2),
                   us_min=round(min(ts), 2), gbs=round(nbytes / med / 1e3, 1), frac_of_6551=round(nbytes / med / 1e3 / 6551, 4))
        fout.write(json.dumps(rec) + "\n"); fout.flush()
        print("%-34s %-8s %-5s %-28s %-4s med %8.1f us  min %8.1f us  frac %.3f" % (wl.name, args.regime, args.dtype, name, op, med, min(ts), rec["frac_of_6551"]), flush=True)
    for k in opts:
        ocpg_b200.set_option(k, 0)
