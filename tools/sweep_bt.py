#!/usr/bin/env python
"""BASELINE.json configs[4]: the b*t sweep.  N total frames in {8 .. 512}, sharded N/G per GPU over G = WORLD_SIZE GPUs
(strong scaling: the total work is fixed as G grows), 6-layer encoder forward + backward with the NCCL all-reduce of the
replicated weights' gradients (7 693 056 fp32 = 30.8 MB, one bucket per layer, issued from backward hooks of the last
micro-batch), for the A2D (S = 4820) and the YTVOS (S = 15 300) geometry.

One process per GPU; ONE launch per G runs the whole sweep (process start-up and NCCL init dominate a single point):

    python tools/sweep_bt.py                                                                  # G = 1
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 tools/sweep_bt.py

Every point is measured twice -- with the all-reduce and with it switched off -- so that the EXPOSED all-reduce time per
step (what the overlap with the backward does not hide) is reported next to the step time; the max over ranks of the
device-timed step is what counts.  Rank 0 appends one JSON line per point to --out.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", default="8,16,32,64,128,256,512")
    ap.add_argument("--shapes", default="a2d,ytvos")
    ap.add_argument("--gemm", default="tf32", choices=["fp32", "tf32"])
    ap.add_argument("--micro", type=int, default=16)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--layers", type=int, default=6)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "sweep_bt.jsonl"))
    args = ap.parse_args()

    import ocpg_b200
    from ocpg_b200 import dist as D
    from ocpg_b200.encoder import build_encoder
    from ocpg_b200.workloads import encoder_workload, shard_frames

    rank, local_rank, world = D.init("nccl")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    ocpg_b200.lib()
    ocpg_b200.set_strict(True)                 # no helper may fall back to torch unnoticed
    torch.backends.cuda.matmul.allow_tf32 = args.gemm == "tf32"
    torch.backends.cudnn.allow_tf32 = args.gemm == "tf32"
    hw = {"a2d": (360, 640), "ytvos": (640, 1152)}

    torch.manual_seed(0)
    enc = build_encoder(num_layers=args.layers, d_ffn=2048, dropout=0.0, fused=True).to(dev)
    enc.train()
    with torch.no_grad():
        for layer in enc.layers:
            layer.self_attn.sampling_offsets.weight.normal_(0, 0.01)
            layer.self_attn.attention_weights.weight.normal_(0, 0.02)
    reducer = D.BucketedGradAllReduce([layer.parameters() for layer in enc.layers])
    n_params = sum(p.numel() for p in enc.parameters())
    fout = open(args.out, "a") if rank == 0 else None

    for shape in args.shapes.split(","):
        for n_total in (int(x) for x in args.frames.split(",")):
            first, frames = shard_frames(n_total, world, rank)
            wl = encoder_workload(f"encoder_{shape}", max(frames, 1), *hw[shape])
            S = wl.S
            shapes_t = torch.tensor(wl.levels, dtype=torch.int64, device=dev)
            start = torch.cat((shapes_t.new_zeros(1), shapes_t.prod(1).cumsum(0)[:-1]))
            g = torch.Generator(device=dev).manual_seed(1000 + first)
            src = torch.randn(frames, S, 256, device=dev, generator=g)
            pos = torch.randn(frames, S, 256, device=dev, generator=g) * 0.1
            vr = torch.ones(frames, wl.L, 2, device=dev)
            gseed = torch.randn(frames, S, 256, device=dev, generator=g)
            micro = max(1, min(args.micro, frames))
            chunks = [(i, min(i + micro, frames)) for i in range(0, frames, micro)]

            def step(allreduce):
                for p in enc.parameters():
                    p.grad = None
                for k, (a, b) in enumerate(chunks):
                    reducer.enabled = allreduce and k == len(chunks) - 1
                    x = src[a:b].clone().requires_grad_(True)
                    out = enc(x, shapes_t, start, vr[a:b], pos[a:b], None)
                    out.backward(gseed[a:b])
                return reducer.finish()

            res = {}
            for allreduce in (True, False):
                for _ in range(args.warmup):
                    step(allreduce)
                torch.cuda.synchronize(); D.barrier(); torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                nbytes = 0
                for _ in range(args.steps):
                    nbytes = step(allreduce)
                e1.record()
                torch.cuda.synchronize(); D.barrier(); torch.cuda.synchronize()
                res[allreduce] = (D.max_over_ranks(e0.elapsed_time(e1) / args.steps, device=dev), nbytes)
            if rank == 0:
                ms_ar, nbytes = res[True]
                ms_no = res[False][0]
                line = {"config": "configs[4] b*t sweep", "shape": shape, "S": S, "total_frames": n_total, "n_gpus": world,
                        "frames_per_gpu": frames, "micro_batch_frames": micro, "layers": args.layers, "gemm_policy": args.gemm,
                        "ms_per_step": round(ms_ar, 3), "ms_per_step_without_allreduce": round(ms_no, 3),
                        "exposed_allreduce_ms": round(ms_ar - ms_no, 3), "allreduce_bytes_per_step": nbytes,
                        "weight_gradients": n_params, "queries_per_s": n_total * S / (ms_ar * 1e-3), "scaling": "strong"}
                fout.write(json.dumps(line) + "\n"); fout.flush()
                print(json.dumps(line), flush=True)
            del src, pos, gseed
            torch.cuda.empty_cache()
    D.finalize()


if __name__ == "__main__":
    main()
