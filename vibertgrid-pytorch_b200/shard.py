"""Document sharding across ranks (one process per GPU) for the joint forward.

The forward has no cross-document term: rank r of W simply owns documents r, r+W, r+2W, ... of the
stream -- the same assignment the reference's ``DistributedSampler(shuffle=False)`` makes
(reference data/SROIE_dataset.py:314-318) -- and no data-path collective exists.  The only collectives are the
bookkeeping ones below: MAX over ranks of the timed region, SUM of the documents processed.

The training step adds the one real exchange of the path: the gradient average over ranks after ``loss.backward()``
(``allreduce_gradients`` -- what DistributedDataParallel does for the reference, train_SROIE.py:203-207), NCCL over NVLink.
"""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.distributed as dist


def shard_documents(n_docs: int, rank: int, world: int) -> List[int]:
    """Indices of the documents rank ``rank`` owns (round-robin, like DistributedSampler without shuffle/padding)."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    return list(range(rank, n_docs, world))


def batch_seed(rank: int, world: int, step: int, rotation: int) -> int:
    """Seed of the synthetic batch a rank runs at a step: distinct across ranks, periodic in ``rotation``."""
    return 1000 * rank + (step % rotation)


def aggregate_throughput(ms_local: float, docs_local: int, device=None) -> Tuple[float, int]:
    """(max over ranks of the timed milliseconds, total documents over ranks).  Works on NCCL (GPU tensors) and
    gloo (CPU tensors); with no process group it is the identity."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(ms_local), int(docs_local)
    t = torch.tensor([float(ms_local)], dtype=torch.float64, device=device)
    n = torch.tensor([int(docs_local)], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(n, op=dist.ReduceOp.SUM)
    return float(t.item()), int(n.item())


def allreduce_gradients(params, bucket_bytes: int = 128 << 20) -> int:
    """Average ``.grad`` of ``params`` over the ranks of the default process group, in flat buckets of about ``bucket_bytes``
    (few large NCCL all-reduces: NVSwitch bandwidth, not launch latency).  Parameters without a gradient are skipped -- every
    rank skips the same ones (the BERT pooler and the torchvision ``fc`` never receive one).  Returns the number of buckets."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return 0
    world = dist.get_world_size()
    grads = [p.grad for p in params if p.grad is not None]
    buckets, cur, size = [], [], 0
    for g in grads:
        if cur and (size + g.numel() * g.element_size() > bucket_bytes or g.dtype != cur[0].dtype):
            buckets.append(cur)
            cur, size = [], 0
        cur.append(g)
        size += g.numel() * g.element_size()
    if cur:
        buckets.append(cur)
    for b in buckets:
        flat = torch._utils._flatten_dense_tensors(b)
        dist.all_reduce(flat)
        flat.div_(world)
        for g, f in zip(b, torch._utils._unflatten_dense_tensors(flat, b)):
            g.copy_(f)
    return len(buckets)


def allreduce_step_arena(net, chunks: int = None) -> int:
    """Gradient average of a GRAPHED training step (train_engine.TrainEngine._graphed_loss): the replayed CUDA graph left every
    parameter gradient in ONE flat fp32 arena (gradients are views of it), so the exchange is ``chunks`` (default 2) in-place NCCL
    all-reduces (op = AVG: the division happens inside the collective) over slices of that arena -- no flatten, no unflatten,
    no separate division pass.  Call it between ``loss = net(batch)`` and ``loss.backward()``: the backward of the step's one
    autograd node hands the (now averaged) arena views to the parameters.  Returns the bytes reduced (0: nothing to do --
    not a graphed step, or a single rank; use ``allreduce_gradients`` after ``backward()`` then)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return 0
    eng = getattr(net, "_train_engine", None)
    arena = getattr(eng, "grad_arena", None) if eng is not None else None
    if arena is None or not getattr(eng, "arena_fresh", False):
        return 0
    eng.arena_fresh = False
    if chunks is None:
        import os
        chunks = int(os.environ.get("VBG_ARENA_CHUNKS", "2"))      # measured at 8 ranks: 1 / 2 / 4 slices -> 2.25 / 2.15 / 2.40 ms
    n = arena.numel()
    step = -(-n // max(1, chunks))
    step = -(-step // 1024) * 1024
    for off in range(0, n, step):
        dist.all_reduce(arena[off:min(n, off + step)], op=dist.ReduceOp.AVG)
    return n * arena.element_size()
