import dataclasses, json, os, statistics, sys
sys.path.insert(0, "/root/repo")
import torch, ocpg_b200
import ocpg_b200.MultiScaleDeformableAttention as MSDA
from ocpg_b200.workloads import A2D_ENCODER, YTVOS_ENCODER, make_inputs
dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for base, ns in ((A2D_ENCODER, (1, 2, 3)), (YTVOS_ENCODER, (1,))):
    for n in ns:
        wl = dataclasses.replace(base, n_frames=n)
        x = make_inputs(wl, "init", seed=0, device=dev)
        for algo in (0, 1):
            ocpg_b200.set_option("bwd_algo", algo)
            fn = lambda: MSDA.ms_deform_attn_backward(x["value"], x["shapes"], x["start"], x["loc"], x["attn"], x["grad_out"], 64)
            for _ in range(3): fn()
            ts = []
            for _ in range(15):
                flush.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); fn(); b.record(); torch.cuda.synchronize()
                ts.append(a.elapsed_time(b) * 1e3)
            print(wl.name, n, "sorted" if algo == 0 else "tiled", round(statistics.median(ts), 1), flush=True)
ocpg_b200.set_option("bwd_algo", 0)
