#!/usr/bin/env python
"""Encoder-level benchmark: BASELINE.json configs[2] (YTVOS shape, full 6-layer encoder fwd+bwd) and configs[4]
(b*t sweep sharded over 1/2/4/8 GPUs with an NCCL all-reduce of the projection-weight gradients).

    python tools/bench_encoder.py --shape ytvos --frames-per-gpu 10
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/bench_encoder.py --shape a2d --frames-per-gpu 64 --micro 16

One step = forward + backward of the 6-layer DeformableTransformerEncoder (ocpg_b200/encoder.py, the re-hosted
deformable_transformer.py:220-290) over this rank's frames, in micro-batches with gradient accumulation (frames are
independent), then one all-reduce per encoder layer of the replicated weights' gradients, issued
from backward hooks of the LAST micro-batch so that it overlaps the remaining backward.  Weak scaling: every rank
processes --frames-per-gpu frames.  The encoder is GEMM-dominated (2.6 MFLOP/query/layer forward), so every number is
quoted with its GEMM precision policy (--gemm fp32 = the reference's: autocast disabled, allow_tf32 False; tf32;
bf16 = autocast for the FFN / projections, the op itself stays fp32), and the operator's own share of the step is
timed inside the run with CUDA events around each library call.

Prints one JSON line on rank 0.
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


def parse_args(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="a2d", choices=["a2d", "ytvos", "t384", "t448", "t512"])
    ap.add_argument("--frames-per-gpu", type=int, default=0, help="default: 5 (a2d, t*) / 10 (ytvos)")
    ap.add_argument("--micro", type=int, default=0, help="frames per micro-batch (default: all, at most 64)")
    ap.add_argument("--layers", type=int, default=6)
    ap.add_argument("--d-ffn", type=int, default=2048)
    ap.add_argument("--gemm", default="fp32", choices=["fp32", "tf32", "bf16"])
    ap.add_argument("--unfused", action="store_true", help="the reference's exact module graph (softmax etc. in torch)")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--no-op-timing", action="store_true")
    ap.add_argument("--graph", action="store_true", help="capture the whole step (all micro-batches, forward + backward) in ONE "
                    "CUDA graph and replay it: no host launch latency.  Single GPU (the all-reduce hooks stay eager)")
    ap.add_argument("--dropout", type=float, default=0.0, help="dropout probability of the layers (reference training: 0.1)")
    return ap.parse_args(argv)


def shape_workload(name, frames):
    from ocpg_b200.workloads import encoder_workload
    hw = {"a2d": (360, 640), "ytvos": (640, 1152), "t384": (384, 640), "t448": (448, 640), "t512": (512, 640)}[name]
    return encoder_workload(f"encoder_{name}_{hw[0]}x{hw[1]}", frames, *hw)


def main(argv=None):
    args = parse_args(argv)
    import ocpg_b200
    import ocpg_b200.MultiScaleDeformableAttention as MSDA
    from ocpg_b200 import dist as D
    from ocpg_b200.encoder import build_encoder

    rank, local_rank, world = D.init("nccl")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    ocpg_b200.lib()                                          # fail loudly if the extension is missing
    frames = args.frames_per_gpu or (10 if args.shape == "ytvos" else 5)
    micro = min(args.micro or frames, 64, frames)
    wl = shape_workload(args.shape, frames)
    torch.backends.cuda.matmul.allow_tf32 = args.gemm == "tf32"
    torch.backends.cudnn.allow_tf32 = args.gemm == "tf32"

    torch.manual_seed(0)                                     # replicated weights: same seed on every rank
    enc = build_encoder(num_layers=args.layers, d_ffn=args.d_ffn, dropout=args.dropout, fused=not args.unfused).to(dev)
    enc.train()
    with torch.no_grad():                                    # offsets / logits that depend on the query, init-like spread
        for layer in enc.layers:
            layer.self_attn.sampling_offsets.weight.normal_(0, 0.01)
            layer.self_attn.attention_weights.weight.normal_(0, 0.02)
    reducer = D.BucketedGradAllReduce([layer.parameters() for layer in enc.layers])
    n_params = sum(p.numel() for p in enc.parameters())

    g = torch.Generator(device=dev).manual_seed(1000 + rank)  # every rank: its own frames
    S = wl.S
    shapes = torch.tensor(wl.levels, dtype=torch.int64, device=dev)
    start = torch.cat((shapes.new_zeros(1), shapes.prod(1).cumsum(0)[:-1]))
    src = torch.randn(frames, S, 256, device=dev, generator=g)
    pos = torch.randn(frames, S, 256, device=dev, generator=g) * 0.1
    vr = torch.ones(frames, wl.L, 2, device=dev)
    gseed = torch.randn(frames, S, 256, device=dev, generator=g)
    chunks = [(i, min(i + micro, frames)) for i in range(0, frames, micro)]

    def step():
        for p in enc.parameters():
            p.grad = None
        for k, (a, b) in enumerate(chunks):
            reducer.enabled = k == len(chunks) - 1
            x = src[a:b].clone().requires_grad_(True)
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=args.gemm == "bf16"):
                out = enc(x, shapes, start, vr[a:b], pos[a:b], None)
            out.backward(gseed[a:b].to(out.dtype))
        return reducer.finish()

    n0 = ocpg_b200.launch_count()
    eager_step, launches_per_step = step, None
    if args.graph:
        if world > 1:
            raise SystemExit("--graph: single GPU only")
        args.no_op_timing = True                 # CUDA events around library calls cannot be captured
        from ocpg_b200.graph import GraphedStep
        c0 = None

        def counted_step():
            nonlocal c0
            c0 = ocpg_b200.launch_count()
            return eager_step()
        graphed = GraphedStep(counted_step, params=enc.parameters())     # warm-up off the default stream, then one capture
        launches_per_step = ocpg_b200.launch_count() - c0

        def step():
            graphed()
            return 0
    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    D.barrier()
    torch.cuda.synchronize()
    if not args.no_op_timing:
        MSDA.start_timing()
    n1 = ocpg_b200.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    comm_bytes = 0
    for _ in range(args.steps):
        comm_bytes = step()
    e1.record()
    torch.cuda.synchronize()
    D.barrier()
    torch.cuda.synchronize()
    launches = ocpg_b200.launch_count() - n1 if launches_per_step is None else launches_per_step * args.steps
    ms = D.max_over_ranks(e0.elapsed_time(e1) / args.steps, device=dev)
    op = {} if args.no_op_timing else MSDA.stop_timing()
    mem_gb = torch.cuda.max_memory_allocated(dev) / 2**30
    if rank == 0:
        queries = frames * S * world
        fb, bb = wl.algorithmic_bytes(4, 4)
        line = {
            "metric": "deformable_encoder_fwd_bwd_queries_per_sec", "value": queries / (ms * 1e-3), "unit": "queries/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "dtype": "f32 operator; GEMMs " + args.gemm, "data": "synthetic",
            "config": {"workload": wl.name, "layers": args.layers, "d_ffn": args.d_ffn, "frames_per_gpu": frames,
                       "micro_batch_frames": micro, "levels": [list(x) for x in wl.levels], "S": S,
                       "module": "reference graph (unfused)" if args.unfused else "fused softmax/locations, no emitted locations",
                       "gemm_policy": args.gemm, "dropout": args.dropout,
                       "launch": "one CUDA graph per step" if args.graph else "eager", "parallelism": f"frames sharded over {world} GPU(s); NCCL all-reduce of "
                       f"{n_params} weight gradients, one bucket per layer, overlapped with backward"},
            "allreduce_bytes_per_step": comm_bytes, "gpu_launches": launches, "peak_mem_gb": round(mem_gb, 2),
        }
        if op:
            f_n, f_ms = op.get("forward", (0, 0.0))
            b_n, b_ms = op.get("backward", (0, 0.0))
            f_ms, b_ms = f_ms / args.steps, b_ms / args.steps
            calls = max(1, f_n // args.steps)
            line["operator"] = {
                "forward_ms_per_step": f_ms, "backward_ms_per_step": b_ms, "calls_per_step": calls,
                "share_of_step": (f_ms + b_ms) / ms,
                "fwd_gbs": fb * args.layers / (f_ms * 1e-3) / 1e9 if f_ms else None,      # algorithmic bytes / device time
                "bwd_gbs": bb * args.layers / (b_ms * 1e-3) / 1e9 if b_ms else None,
            }
        print(json.dumps(line), flush=True)
    D.finalize()


if __name__ == "__main__":
    main()
