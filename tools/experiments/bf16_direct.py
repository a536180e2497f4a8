#!/usr/bin/env python
"""bf16 value: grad_value accumulated (a) in an fp32 buffer + one conversion pass, (b) by packed bf16 reds straight into
the bf16 tensor (msda_backward_bf16 with a NULL fp32 buffer).  Reports, for the A2D and the YTVOS encoder shapes, the error
of grad_value against the fp64 oracle (fed the bf16-rounded inputs) as max|a-b| / max|b| and as RMS, and the device time
of the whole backward call (zero-fill, kernel, conversion) next to the fp32 backward."""
import json, os, statistics, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch
import ocpg_b200.MultiScaleDeformableAttention as MSDA
from ocpg_b200.workloads import A2D_ENCODER, YTVOS_ENCODER, make_inputs
import oracle

dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return statistics.median(ts)


for wl in (A2D_ENCODER, YTVOS_ENCODER):
    for regime in ("init", "uniform"):
        x = make_inputs(wl, regime, seed=5)
        xb = {k: (v.bfloat16() if k in ("value", "grad_out") else v) for k, v in x.items()}
        n = lambda t: t.float().numpy() if t.is_floating_point() else t.numpy()
        want = oracle.c_backward(n(xb["value"]), n(x["shapes"]), n(x["start"]), n(x["loc"]), n(x["attn"]), n(xb["grad_out"]), np.float64)[0]
        d32 = {k: v.to(dev) for k, v in x.items()}
        d16 = {k: v.to(dev) for k, v in xb.items()}
        call = lambda d: MSDA.ms_deform_attn_backward(d["value"], d["shapes"], d["start"], d["loc"], d["attn"], d["grad_out"], 64)
        rec = dict(workload=wl.name, regime=regime)
        rec["fp32_bwd_us"] = round(timeit(lambda: call(d32)), 1)
        for mode in (False, True):
            MSDA.set_bf16_grad_value_direct(mode)
            gv = call(d16)[0].float().cpu().numpy().astype(np.float64)
            err = np.abs(gv - want)
            tag = "direct_bf16_reds" if mode else "fp32_buffer"
            rec[tag] = dict(bwd_us=round(timeit(lambda: call(d16)), 1), max_err_over_max=float(err.max() / np.abs(want).max()),
                            rms_err_over_rms=float(np.sqrt((err ** 2).mean()) / np.sqrt((want ** 2).mean())))
        MSDA.set_bf16_grad_value_direct(False)
        print(json.dumps(rec), flush=True)
