#!/usr/bin/env python
"""Backward time per frame vs number of frames (A2D geometry): how much of the 5-frame launch is the tail of the last wave
of persistent CTAs (6520 passes over 592 workers = 11.01 waves)?"""
import dataclasses, json, os, statistics, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import ocpg_b200.MultiScaleDeformableAttention as MSDA
from ocpg_b200.workloads import A2D_ENCODER, make_inputs
dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for n in (1, 2, 3, 4, 5, 6, 8, 10, 16, 20, 40):
    wl = dataclasses.replace(A2D_ENCODER, n_frames=n)
    x = make_inputs(wl, "init", seed=0, device=dev)
    for op, fn in (("fwd", lambda: MSDA.ms_deform_attn_forward(x["value"], x["shapes"], x["start"], x["loc"], x["attn"], 64)),
                   ("bwd", lambda: MSDA.ms_deform_attn_backward(x["value"], x["shapes"], x["start"], x["loc"], x["attn"], x["grad_out"], 64))):
        for _ in range(3):
            fn()
        ts = []
        for _ in range(15):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) * 1e3)
        med = statistics.median(ts)
        print(json.dumps(dict(op=op, frames=n, us=round(med, 1), us_per_frame=round(med / n, 2))), flush=True)
