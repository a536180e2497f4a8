"""Multi-GPU plumbing for the MSDeformAttn hot path: one process per GPU, frames sharded, no data-path
collective (every output row (n, q) reads only frame n, so the flattened
N = b*t axis splits into independent units -- the way the reference scales with DDP over clips,
main.py:62).  torch.distributed is used only for the barrier and the max-over-ranks of timings
(NCCL on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

import os
from typing import Tuple

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


def finalize():
    if dist.is_initialized():
        dist.destroy_process_group()
