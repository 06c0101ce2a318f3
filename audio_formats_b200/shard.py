"""File sharding across GPUs: streams are independent, so they are assigned to ranks by
estimated work with no data-path collective.  Only the timing join uses torch.distributed."""
from __future__ import annotations

from typing import Sequence


def shard_contiguous(n_items: int, rank: int, world: int) -> range:
    """Even contiguous split (used when every stream costs the same, e.g. the homogeneous bench)."""
    base, rem = divmod(n_items, world)
    start = rank * base + min(rank, rem)
    return range(start, start + base + (1 if rank < rem else 0))


def shard_lpt(costs: Sequence[float], world: int) -> list[list[int]]:
    """Longest-processing-time-first greedy assignment by cost (granule-channels of each stream), so
    heterogeneous batches balance across GPUs.  Deterministic: ties go to the lower rank / lower index."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    loads = [0.0] * world
    out: list[list[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (loads[k], k))
        out[r].append(i)
        loads[r] += costs[i]
    for lst in out:
        lst.sort()
    return out
