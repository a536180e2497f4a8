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


def _reducer_worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from ocpg_b200 import dist as D
    D.init("gloo")
    torch.manual_seed(0)                                       # replicated weights
    layers = torch.nn.ModuleList([torch.nn.Linear(6, 6) for _ in range(3)])
    red = D.BucketedGradAllReduce([l.parameters() for l in layers])
    g = torch.Generator().manual_seed(100 + rank)              # each rank: its own frames
    x = torch.randn(4, 6, generator=g)
    # two micro-batches with gradient accumulation: reduce only after the last one
    red.enabled = False
    h = x[:2]
    for l in layers:
        h = torch.tanh(l(h))
    h.sum().backward()
    red.enabled = True
    h = x[2:]
    for l in layers:
        h = torch.tanh(l(h))
    h.sum().backward()
    nbytes = red.finish()
    q.put((rank, [p.grad.tolist() for p in layers.parameters()], x.tolist(), nbytes))     # plain lists: no shared-memory handles
    D.finalize()


def test_bucketed_grad_allreduce_gloo():
    """Per-layer buckets issued from backward hooks == the gradient of the mean over ranks of the summed loss."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_reducer_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted((q.get(timeout=120) for _ in ps), key=lambda t: t[0])
    for p in ps:
        p.join(60)
        assert p.exitcode == 0
    torch.manual_seed(0)
    layers = torch.nn.ModuleList([torch.nn.Linear(6, 6) for _ in range(3)])
    for _, _, x, _ in res:
        h = torch.tensor(x)
        for l in layers:
            h = torch.tanh(l(h))
        (h.sum() / 2).backward()                              # average over the 2 ranks
    want = [p.grad for p in layers.parameters()]
    for rank, grads, _, nbytes in res:
        assert nbytes == sum(p.numel() for p in layers.parameters()) * 4
        for g, w in zip(grads, want):
            assert torch.allclose(torch.tensor(g), w, rtol=1e-5, atol=1e-6), rank
