"""Multi-GPU plumbing for the MSDeformAttn hot path: one process per GPU, frames sharded, no data-path
collective (every output row (n, q) reads only frame n, so the flattened
N = b*t axis splits into independent units -- the way the reference scales with DDP over clips,
main.py:62).  torch.distributed is used only for the barrier and the max-over-ranks of timings
(NCCL on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

import os
from typing import Iterable, List, Sequence, Tuple

import torch
import torch.distributed as dist

from .workloads import shard_frames


def env_world() -> Tuple[int, int, int]:
    """(rank, local_rank, world_size) from the torchrun environment (1 process => (0, 0, 1))."""
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def init(backend: str = "nccl"):
    """Join the job's process group if launched under torchrun with WORLD_SIZE > 1."""
    rank, local_rank, world = env_world()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            kw["device_id"] = torch.device("cuda", local_rank)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, local_rank, world


def bind_to_gpu_cpus(local_rank: int):
    """Pin this process to the CPU cores NVML reports as local to its GPU, so that pinned host buffers are
    first-touched on the GPU's own NUMA node and every rank's host<->device copies use its own socket's memory
    (the e2e path of bench.py is PCIe-bound; without this all ranks share one socket's memory bandwidth).
    Returns the CPU list, or None if NVML or the affinity call is unavailable."""
    try:
        import pynvml
        pynvml.nvmlInit()
        uuid = str(torch.cuda.get_device_properties(local_rank).uuid)
        handle = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(handle, words)
        cpus = [64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1]
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return cpus
    except Exception:
        pass
    return None


def barrier():
    if dist.is_initialized():
        dist.barrier()


def max_over_ranks(x: float, device="cpu") -> float:
    """Timings are reported as the max over ranks (the slowest rank bounds the job)."""
    if not dist.is_initialized():
        return float(x)
    t = torch.tensor([float(x)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(x: float, device="cpu") -> float:
    if not dist.is_initialized():
        return float(x)
    t = torch.tensor([float(x)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def local_frames(n_frames_global: int) -> Tuple[int, int]:
    """This rank's contiguous block of the global frame axis: (first_frame, n_local)."""
    rank, _, world = env_world()
    return shard_frames(n_frames_global, world, rank)


class BucketedGradAllReduce:
    """All-reduce of the REPLICATED weights' gradients -- the only collective of the encoder benchmark
    (BASELINE.json configs[4]; the op itself has none).  Parameters are grouped in buckets (one per encoder
    layer); a bucket's all-reduce is issued from a post-accumulate-grad hook the moment its last gradient is
    ready, so it runs on the communication stream while the backward of the earlier layers is still computing.
    ``finish()`` waits for the buckets and writes the averaged gradients back.

    With gradient accumulation over micro-batches set ``enabled = False`` for all but the last micro-batch.
    Without an initialised process group it does nothing."""

    def __init__(self, buckets: Sequence[Iterable[torch.nn.Parameter]], average: bool = True):
        self.buckets: List[List[torch.nn.Parameter]] = [[p for p in b if p.requires_grad] for b in buckets]
        self.average = average
        self.enabled = True
        self._pending = [len(b) for b in self.buckets]
        self._inflight = []
        self._handles = []
        for bi, bucket in enumerate(self.buckets):
            for p in bucket:
                self._handles.append(p.register_post_accumulate_grad_hook(self._make_hook(bi)))

    def _make_hook(self, bi: int):
        def hook(_param):
            if not self.enabled or not dist.is_initialized():
                return
            self._pending[bi] -= 1
            if self._pending[bi] == 0:
                bucket = self.buckets[bi]
                flat = torch.cat([p.grad.reshape(-1) for p in bucket])
                work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, async_op=True)
                self._inflight.append((bi, flat, work))
        return hook

    def finish(self) -> int:
        """Wait for every issued bucket, scatter the reduced values back into ``.grad``; returns the number of
        gradient bytes this rank contributed to collectives."""
        nbytes = 0
        world = dist.get_world_size() if dist.is_initialized() else 1
        for bi, flat, work in self._inflight:
            work.wait()
            if self.average:
                flat.div_(world)
            off = 0
            for p in self.buckets[bi]:
                n = p.numel()
                p.grad.copy_(flat[off:off + n].view_as(p.grad))
                off += n
            nbytes += flat.numel() * flat.element_size()
        self._inflight.clear()
        self._pending = [len(b) for b in self.buckets]
        return nbytes

    def remove(self):
        for h in self._handles:
            h.remove()
        self._handles.clear()


def finalize():
    if dist.is_initialized():
        dist.destroy_process_group()
