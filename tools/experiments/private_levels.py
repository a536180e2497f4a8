#!/usr/bin/env python
"""EXPERIMENT: backward with the coarse pyramid levels of grad_value accumulated into K private scratch copies (one per
warp, round-robin) and folded into grad_value afterwards -- does spreading the hot rows over K x more L2 sectors pay?
(`msda_set_private_workspace`, not part of the public header.)"""
import ctypes, json, os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import ocpg_b200
import ocpg_b200.MultiScaleDeformableAttention as MSDA
from ocpg_b200.workloads import A2D_ENCODER, YTVOS_ENCODER, make_inputs

L = ocpg_b200.lib()
L.msda_set_private_workspace.argtypes = [ctypes.c_void_p, ctypes.c_uint64] + [ctypes.c_int] * 4
dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for wl, nsets, iters in ((A2D_ENCODER, 4, 20), (YTVOS_ENCODER, 2, 10)):
    for regime in ("init", "uniform"):
        sets = [make_inputs(wl, regime, seed=i, device=dev) for i in range(nsets)]
        bwd = lambda x: MSDA.ms_deform_attn_backward(x["value"], x["shapes"], x["start"], x["loc"], x["attn"], x["grad_out"], 64)
        L.msda_set_private_workspace(None, 0, 0, 0, 0, 0)
        ref = bwd(sets[0])
        starts = [0]
        for h, w in wl.levels:
            starts.append(starts[-1] + h * w)
        for level, copies in ((4, 0), (3, 8), (3, 32), (2, 8), (2, 32), (1, 8), (2, 128)):
            ws = None
            if copies:
                rows = wl.S - starts[level]
                ws = torch.empty(copies * wl.n_frames * rows * 256, dtype=torch.float32, device=dev)
                L.msda_set_private_workspace(ws.data_ptr(), ws.numel() * 4, level, copies, starts[level], rows)
            else:
                L.msda_set_private_workspace(None, 0, 0, 0, 0, 0)
            got = bwd(sets[0])
            err = max(float((a - b).abs().max() / b.abs().max()) for a, b in zip(got, ref))
            for i in range(3):
                bwd(sets[i % nsets])
            torch.cuda.synchronize()
            ts = []
            for i in range(iters):
                flush.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); bwd(sets[i % nsets]); b.record()
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b) * 1e3)
            print(json.dumps(dict(workload=wl.name, regime=regime, first_private_level=level, copies=copies,
                                  ws_mb=0 if ws is None else round(ws.numel() * 4 / 1e6, 1), us_median=round(statistics.median(ts), 1),
                                  us_min=round(min(ts), 1), max_rel_diff_vs_plain=err)), flush=True)
            del ws
L.msda_set_private_workspace(None, 0, 0, 0, 0, 0)
