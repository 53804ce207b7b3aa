"""Utterance sharding across the GPUs of one box.

The decode path has no cross-utterance term, so multi-GPU is plain data parallelism: every rank owns
a full weight replica, a private KV pool and a contiguous slice of the utterances; there is no
collective on the decode path.  ``torch.distributed`` is used only to gather the emitted codes on the
host and to reduce timings (max over ranks).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch


def partition(n_items: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous, balanced [start, end) ranges; the first ``n_items % world`` ranks get one more."""
    base, extra = divmod(n_items, world)
    out, start = [], 0
    for r in range(world):
        n = base + (1 if r < extra else 0)
        out.append((start, start + n))
        start += n
    return out


def shard(items: Sequence, rank: int, world: int) -> Tuple[list, List[int]]:
    """This rank's slice of ``items`` and the global ids of its elements (the RNG's ``seq_id``)."""
    lo, hi = partition(len(items), world)[rank]
    return list(items[lo:hi]), list(range(lo, hi))


def gather_utterances(local: Sequence[torch.Tensor], dist=None, dst: int = 0) -> Optional[List[torch.Tensor]]:
    """Host-side gather of per-utterance code tensors (ragged) onto ``dst`` in global utterance order."""
    local = [t.cpu() for t in local]
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    bucket = [None] * world if rank == dst else None
    dist.gather_object(local, bucket, dst=dst)
    if rank != dst:
        return None
    return [t for part in bucket for t in part]


def max_over_ranks(values: Sequence[float], dist=None, device=None) -> List[float]:
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t.tolist()]
