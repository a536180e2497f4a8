#!/usr/bin/env python
"""Decoder-level benchmark: BASELINE.json configs[3] at layer level -- the decoder stack (cross-attention of 5 object
queries per frame over the full multi-level memory, self-attention, FFN, the reference-point / top-30 consumers) forward +
backward over the memory of an A2D clip (5 frames, S = 4820).

    python tools/bench_decoder.py [--layers 4] [--graph] [--unfused] [--dropout 0.1]

One step = forward + backward of ocpg_b200.decoder.DeformableTransformerDecoder (the re-hosted
deformable_transformer.py:293-398) with gradients into tgt, the memory and every weight.  With 25 queries the stack is
pure launch latency: the numbers to read are launches per step and the eager / CUDA-graph times.  ``--unfused`` runs the
reference's module graph (torch softmax / location arithmetic / LayerNorms / topk) on the same operator.
Prints one JSON line.
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


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--layers", type=int, default=4, help="opts.py: dec_layers default 4")
    ap.add_argument("--d-ffn", type=int, default=2048)
    ap.add_argument("--queries", type=int, default=5, help="object queries per frame (opts.py:64)")
    ap.add_argument("--frames", type=int, default=5)
    ap.add_argument("--dropout", type=float, default=0.0)
    ap.add_argument("--unfused", action="store_true")
    ap.add_argument("--graph", action="store_true")
    ap.add_argument("--gemm", default="tf32", choices=["fp32", "tf32"])
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    args = ap.parse_args(argv)

    import ocpg_b200
    from ocpg_b200 import decoder as D
    from ocpg_b200.workloads import A2D_ENCODER as wl
    dev = torch.device("cuda:0")
    ocpg_b200.lib()
    torch.backends.cuda.matmul.allow_tf32 = args.gemm == "tf32"
    torch.manual_seed(0)
    dec = D.build_decoder(num_layers=args.layers, d_ffn=args.d_ffn, dropout=args.dropout, fused=not args.unfused).to(dev)
    dec.train()
    with torch.no_grad():
        for layer in dec.layers:
            layer.cross_attn.sampling_offsets.weight.normal_(0, 0.01)
            layer.cross_attn.attention_weights.weight.normal_(0, 0.05)
    if args.unfused:                       # the reference's consumers as well
        D._native_ok = lambda *t: False
    N, Lq, S = args.frames, args.queries, wl.S
    g = torch.Generator(device=dev).manual_seed(3)
    shapes = torch.tensor(wl.levels, dtype=torch.int64, device=dev)
    start = torch.cat((shapes.new_zeros(1), shapes.prod(1).cumsum(0)[:-1]))
    memory = torch.randn(N, S, 256, device=dev, generator=g, requires_grad=True)
    tgt = torch.randn(N, Lq, 256, device=dev, generator=g, requires_grad=True)
    qpos = torch.randn(N, Lq, 256, device=dev, generator=g)
    ref = (0.2 + 0.6 * torch.rand(N, Lq, 2, device=dev, generator=g)).requires_grad_(True)
    vr = torch.ones(N, wl.L, 2, device=dev)
    ghs = torch.randn(args.layers, N, Lq, 256, device=dev, generator=g)

    def eager_step():
        for p in dec.parameters():
            p.grad = None
        memory.grad = tgt.grad = ref.grad = None
        hs, refs, samples = dec(tgt, ref, memory, shapes, start, vr, qpos, None)
        hs.backward(ghs)
        return samples

    step, launches_per_step = eager_step, None
    if args.graph:
        from ocpg_b200.graph import GraphedStep
        c0 = [0]

        def counted_step():
            c0[0] = ocpg_b200.launch_count()
            return eager_step()
        graphed = GraphedStep(counted_step, params=dec.parameters())
        launches_per_step = ocpg_b200.launch_count() - c0[0]
        step = graphed
    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    n1 = ocpg_b200.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    launches = launches_per_step if launches_per_step is not None else (ocpg_b200.launch_count() - n1) // args.steps
    print(json.dumps({
        "metric": "deformable_decoder_fwd_bwd_ms_per_step", "value": ms, "unit": "ms", "higher_is_better": False,
        "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "data": "synthetic", "dtype": "f32 operator; GEMMs " + args.gemm,
        "queries_per_sec": N * Lq / (ms * 1e-3),
        "config": {"workload": f"a2d_decoder_{args.layers}layers_N{N}_Lq{Lq}_S{S}", "d_ffn": args.d_ffn, "dropout": args.dropout,
                   "module": "reference graph (unfused)" if args.unfused else "fused operator + epilogue + consumer kernels",
                   "launch": "one CUDA graph per step" if args.graph else "eager"},
        "library_launches_per_step": launches}), flush=True)


if __name__ == "__main__":
    main()
