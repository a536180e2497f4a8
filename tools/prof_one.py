#!/usr/bin/env python
"""A short fixed sequence of forward / backward launches for ncu (A2D encoder shape by default)."""
import argparse
import os
import sys

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
ap.add_argument("--reps", type=int, default=4)
ap.add_argument("--opt", action="append", default=[], help="key=value library option")
args = ap.parse_args()
for o in args.opt:
    k, v = o.split("=")
    ocpg_b200.set_option(k, int(v))
wl = {"a2d": A2D_ENCODER, "ytvos": YTVOS_ENCODER, "decoder": A2D_DECODER}[args.workload]
dev = torch.device("cuda:0")
vdt = torch.bfloat16 if args.dtype == "bf16" else None
sets = [make_inputs(wl, args.regime, seed=i, device=dev, value_dtype=vdt) for i in range(2)]
torch.cuda.synchronize()
for i in range(args.reps):
    x = sets[i % 2]
    out = MSDA.ms_deform_attn_forward(x["value"], x["shapes"], x["start"], x["loc"], x["attn"], 64)
    g = MSDA.ms_deform_attn_backward(x["value"], x["shapes"], x["start"], x["loc"], x["attn"], x["grad_out"], 64)
torch.cuda.synchronize()
print("ok", float(out.float().abs().sum()), float(g[0].float().abs().sum()))
