"""world_size-2 gloo test of the N>1 host path (frame sharding + max-over-ranks timing plumbing)."""
import os
import socket

import torch
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from ocpg_b200 import dist as D
    from ocpg_b200.workloads import encoder_workload, make_inputs
    D.init("gloo")
    first, n_local = D.local_frames(5)
    wl = encoder_workload("t", 5, 64, 64)
    x = make_inputs(wl, "init", seed=3)                       # every rank builds the same global batch ...
    mine = x["value"][first:first + n_local]                  # ... and keeps its own frames
    total = D.sum_over_ranks(float(mine.sum()))
    slow = D.max_over_ranks(10.0 + rank)
    D.barrier()
    q.put((rank, first, n_local, total, float(x["value"].sum()), slow))
    D.finalize()


def test_two_rank_sharding_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(60)
        assert p.exitcode == 0
    (r0, f0, n0, tot0, full0, slow0), (r1, f1, n1, tot1, full1, slow1) = res
    assert (f0, n0, f1, n1) == (0, 3, 3, 2)                  # contiguous block split, remainder to low ranks
    assert abs(tot0 - full0) < 1e-3 * abs(full0) + 1e-3 and tot0 == tot1   # shards cover the batch exactly once
    assert slow0 == slow1 == 11.0                            # max over ranks
