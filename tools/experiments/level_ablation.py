#!/usr/bin/env python
"""Which pyramid level's reds cost what: backward (full / scatter-only) with the attention weights of some levels
zeroed -- zero-weight corners send no red (msda_sm100.cu: red_row predicate).  Needs a library built with
-DMSDA_EXPERIMENTS (the measurement switch `bwd_mode` is compiled out of the product) and run with MSDA_LIB=<that .so>; it
measures the query-major kernel (`bwd_algo = 1`), which is what the round-1 numbers in profiles/ are of."""
import os, sys, statistics, json
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import ocpg_b200
import ocpg_b200.MultiScaleDeformableAttention as MSDA
from ocpg_b200.workloads import A2D_ENCODER, make_inputs
dev = torch.device("cuda:0")
sets = [make_inputs(A2D_ENCODER, "init", seed=i, device=dev) for i in range(3)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timeit(xs):
    ts = []
    for i in range(15):
        x = xs[i % len(xs)]
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); MSDA.ms_deform_attn_backward(x["value"], x["shapes"], x["start"], x["loc"], x["attn"], x["grad_out"], 64); b.record()
        torch.cuda.synchronize(); ts.append(a.elapsed_time(b) * 1e3)
    return statistics.median(ts[3:])
ocpg_b200.set_option("bwd_algo", 1)
for mode in (0, 2, 1):
    ocpg_b200.set_option("bwd_mode", mode)
    for keep in ([0, 1, 2, 3], [0, 1, 2], [0, 1], [0], [1, 2, 3], [3], [2], [1]):
        xs = []
        for x in sets:
            y = dict(x); a = x["attn"].clone()
            for l in range(4):
                if l not in keep: a[:, :, :, l] = 0
            y["attn"] = a; xs.append(y)
        nz = float((xs[0]["attn"] != 0).float().mean())
        print(json.dumps(dict(bwd_mode=mode, levels_with_reds=keep, us=round(timeit(xs), 1))), flush=True)
ocpg_b200.set_option("bwd_mode", 0)
ocpg_b200.set_option("bwd_algo", 0)
