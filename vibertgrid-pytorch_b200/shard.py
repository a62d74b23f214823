"""Document sharding across ranks (one process per GPU) for the joint forward.

The forward has no cross-document term: rank r of W simply owns documents r, r+W, r+2W, ... of the
stream -- the same assignment the reference's ``DistributedSampler(shuffle=False)`` makes
(reference data/SROIE_dataset.py:314-318) -- and no data-path collective exists.  The only collectives are the
bookkeeping ones below: MAX over ranks of the timed region, SUM of the documents processed.
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
