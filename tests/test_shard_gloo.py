"""N>1 host logic on CPU: world_size-2 gloo process group (127.0.0.1 rendezvous)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vibertgrid_pytorch_b200 import shard


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        docs = shard.shard_documents(11, rank, world)
        ms, total = shard.aggregate_throughput(10.0 + 5.0 * rank, len(docs))
        seeds = [shard.batch_seed(rank, world, s, 4) for s in range(6)]
        # gradient averaging of the training step: same parameter list on both ranks, one parameter without a gradient
        torch.manual_seed(0)
        ps = [torch.nn.Parameter(torch.zeros(n)) for n in (5, 70000, 3, 9)]
        for i, p in enumerate(ps):
            if i != 2:
                p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
        nb = shard.allreduce_gradients(ps, bucket_bytes=100000)
        avg = [None if p.grad is None else float(p.grad.mean()) for p in ps]
        q.put((rank, docs, ms, total, seeds, nb, avg))
    finally:
        dist.destroy_process_group()


def test_world2_sharding_and_max_over_ranks():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, d0, ms0, n0, s0, nb0, a0), (r1, d1, ms1, n1, s1, nb1, a1) = res
    assert nb0 == nb1 and nb0 >= 2 and a0 == a1 == [1.5, 3.0, None, 6.0]     # mean over ranks of (rank + 1) * (i + 1)
    assert sorted(d0 + d1) == list(range(11)) and not set(d0) & set(d1)      # a partition: every document exactly once
    assert ms0 == ms1 == 15.0                                                # MAX over ranks, identical on both
    assert n0 == n1 == 11                                                    # whole-job document count
    assert not set(s0) & set(s1) and s0[0] == s0[4]                          # ranks see different documents; rotation of 4


def test_single_process_is_identity():
    assert shard.aggregate_throughput(3.5, 8) == (3.5, 8)
    assert shard.shard_documents(5, 0, 1) == [0, 1, 2, 3, 4]
