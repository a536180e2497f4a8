"""Whole-step CUDA-graph capture for the re-hosted modules.

At the reference's shapes the layers around the operator are launch-bound: the A2D encoder step issues ~1 500 kernels for
7 ms of device work, the decoder stack ~600 for 1.5 ms.  Nothing in ocpg_b200's forward or backward
synchronises with the host (shape checks and host copies of ``spatial_shapes`` are cached per tensor; dropout reads its
key words from device memory), so a complete training step -- forward, backward, gradient accumulation over micro-batches
-- can be captured once and replayed:

    step = GraphedStep(lambda: loss_fn(model(*static_inputs)).backward(), params=model.parameters())
    for batch in loader:
        copy_into(static_inputs, batch)      # inputs live in fixed buffers
        step()                               # gradients appear in p.grad, as after an eager step
        optimizer.step()

Rules (CUDA graphs'): fixed shapes, inputs updated in place, no host-side reads inside the step.
"""
from __future__ import annotations

from typing import Callable, Iterable, Optional

import torch


class GraphedStep:
    def __init__(self, fn: Callable[[], object], params: Optional[Iterable[torch.nn.Parameter]] = None, warmup: int = 3):
        """``fn`` runs one step eagerly (forward + backward).  It is warmed up ``warmup`` times on a side stream (allocator
        pools, cuBLAS workspaces and the per-tensor caches fill), the gradients of ``params`` are dropped so that the
        capture's backward allocates them inside the graph's pool, and ``fn`` is captured once."""
        self.params = list(params) if params is not None else []
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                fn()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        for p in self.params:
            p.grad = None
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.result = fn()

    def __call__(self):
        self.graph.replay()
        return self.result

    replay = __call__
