#!/usr/bin/env python
"""bwd_pdl = 1: zero-fill of grad_value by msda_zero_fill with the row-major backward as its programmatic dependent.
Checks results against bwd_pdl = 0 (eager and inside a CUDA graph) and times both.
ARCHIVED: the `bwd_pdl` option was measured (slower, tools/experiments/README.md) and is not in the library; this script
documents how it was checked."""
import json, os, statistics, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import ocpg_b200
import ocpg_b200.MultiScaleDeformableAttention as MSDA
from ocpg_b200.workloads import A2D_ENCODER, YTVOS_ENCODER, make_inputs
dev = torch.device("cuda:0")
for wl in (A2D_ENCODER, YTVOS_ENCODER):
    sets = [make_inputs(wl, "init", seed=i, device=dev) for i in range(3)]
    call = lambda x: MSDA.ms_deform_attn_backward(x["value"], x["shapes"], x["start"], x["loc"], x["attn"], x["grad_out"], 64)
    ocpg_b200.set_option("bwd_pdl", 0)
    ref = [call(x) for x in sets]
    res = {}
    for pdl in (0, 1):
        ocpg_b200.set_option("bwd_pdl", pdl)
        for x, r in zip(sets, ref):
            for _ in range(3):
                g = call(x)
                assert torch.equal(g[1], r[1]) and torch.equal(g[2], r[2])
                assert (g[0] - r[0]).abs().max().item() <= 1e-5 * r[0].abs().max().item()
        # CUDA graph: forward + backward per input set, replayed
        graphs, keep = [], []
        side = torch.cuda.Stream()
        with torch.cuda.stream(side):
            call(sets[0])
        torch.cuda.synchronize()
        for x in sets:
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                keep.append(call(x))
            graphs.append(gr)
        for _ in range(2):
            for gr in graphs:
                gr.replay()
        torch.cuda.synchronize()
        for k, r in zip(keep, ref):
            assert torch.equal(k[1], r[1]) and torch.equal(k[2], r[2])
            assert (k[0] - r[0]).abs().max().item() <= 1e-5 * r[0].abs().max().item()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(60):
            graphs[i % 3].replay()
        b.record(); torch.cuda.synchronize()
        res[pdl] = a.elapsed_time(b) / 60 * 1e3
    ocpg_b200.set_option("bwd_pdl", 0)
    print(json.dumps(dict(workload=wl.name, graph_us_pdl0=round(res[0], 1), graph_us_pdl1=round(res[1], 1))), flush=True)
print("ok")
