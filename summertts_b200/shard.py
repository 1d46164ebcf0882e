"""Utterance sharding across replicas (one engine per GPU, no data-path collective).

The path shards by independent units: utterances never exchange state, so a batch is
partitioned across the ranks of one box and each rank runs its own engine.  The only cross-rank
traffic is the timing/throughput reduction done by the caller (a 2-float all-reduce).
"""
from __future__ import annotations

from typing import List, Sequence


def round_robin(n_items: int, world: int, rank: int) -> List[int]:
    """Indices of the utterances rank `rank` owns (strong scaling: fixed global batch)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    return list(range(rank, n_items, world))


def balanced(lengths: Sequence[int], world: int) -> List[List[int]]:
    """Longest-processing-time-first partition by phoneme count (frames scale with ids): keeps the
    replicas' decoder work even when utterance lengths are ragged."""
    order = sorted(range(len(lengths)), key=lambda i: (-lengths[i], i))
    loads = [0] * world
    parts: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (loads[k], k))
        parts[r].append(i)
        loads[r] += lengths[i]
    for p in parts:
        p.sort()
    return parts


def weak_batch(per_gpu: int, world: int, rank: int) -> List[int]:
    """Global utterance ids of rank `rank` under weak scaling (per-GPU work fixed)."""
    return list(range(rank * per_gpu, (rank + 1) * per_gpu))
