#!/usr/bin/env python
"""Development check: resident kernels vs the tiled kernels (which are parity-tested) -- max differences and timings."""
import argparse, json, os, statistics, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import ocpg_b200  # noqa: E402
import ocpg_b200.MultiScaleDeformableAttention as MSDA  # noqa: E402
from ocpg_b200.workloads import A2D_ENCODER, YTVOS_ENCODER, encoder_workload, make_inputs  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--ops", default="fwd,bwd")
ap.add_argument("--cases", default="a2d:init:f32,a2d:uniform:f32,ytvos:init:f32,a2d:init:bf16,t512:init:f32")
ap.add_argument("--variants", default="tiled:resident=-1;res16:resident=1;res8:resident=1,fwd_warps=8")
ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "check_resident.jsonl"))
args = ap.parse_args()
dev = torch.device("cuda:0")
WL = {"a2d": A2D_ENCODER, "ytvos": YTVOS_ENCODER, "t512": encoder_workload("a2d_train_512x640_N6", 6, 512, 640),
      "t448": encoder_workload("a2d_train_448x640_N5", 5, 448, 640)}
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
fout = open(args.out, "a")

def run(op, x):
    if op == "fwd":
        return (MSDA.ms_deform_attn_forward(x["value"], x["shapes"], x["start"], x["loc"], x["attn"], 64),)
    return tuple(MSDA.ms_deform_attn_backward(x["value"], x["shapes"], x["start"], x["loc"], x["attn"], x["grad_out"], 64))

for case in args.cases.split(","):
    wname, regime, dt = case.split(":")
    wl = WL[wname]
    vdt = torch.bfloat16 if dt == "bf16" else None
    nsets = 2 if wname == "ytvos" else 3
    sets = [make_inputs(wl, regime, seed=i, device=dev, value_dtype=vdt) for i in range(nsets)]
    vb = 2 if vdt is not None else 4
    nbytes = dict(zip(("fwd", "bwd"), wl.algorithmic_bytes(vb, vb)))
    base = {}
    for var in args.variants.split(";"):
        name, _, optstr = var.partition(":")
        opts = dict((kv.split("=")[0], int(kv.split("=")[1])) for kv in optstr.split(",") if kv)
        for k, v in opts.items():
            ocpg_b200.set_option(k, v)
        for op in args.ops.split(","):
            res = run(op, sets[0])
            torch.cuda.synchronize()
            diffs = None
            if op not in base:
                base[op] = [r.float().clone() for r in res]
            else:
                diffs = [float((r.float() - b).abs().max() / b.abs().max().clamp_min(1e-30)) for r, b in zip(res, base[op])]
            for i in range(3):
                run(op, sets[i % nsets])
            torch.cuda.synchronize()
            ts = []
            for i in range(args.iters):
                flush.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); run(op, sets[i % nsets]); b.record()
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b) * 1e3)
            med = statistics.median(ts)
            rec = dict(case=case, variant=name, op=op, us_median=round(med, 2), us_min=round(min(ts), 2),
                       frac_of_6551=round(nbytes[op] / med / 1e3 / 6551, 4), rel_diff_vs_first=diffs)
            fout.write(json.dumps(rec) + "\n"); fout.flush()
            print(json.dumps(rec), flush=True)
        for k in opts:
            ocpg_b200.set_option(k, 0)
