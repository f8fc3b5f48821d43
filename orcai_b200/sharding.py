"""Sharding of a recording table across GPUs / ranks.

The path shards by recording: every table row is independent (own global max, own percentiles, own
output file), so ranks never exchange data on the data path; the only communication is a host-side
gather of per-recording results (label rows, errors).  Assignment is longest-processing-time-first
so that one long recording does not serialise the tail.
"""

from __future__ import annotations

import os


def assign_rows(costs: list[float], world: int) -> list[list[int]]:
    """LPT: rows sorted by decreasing cost, each given to the currently least-loaded rank.

    Deterministic (ties broken by row index, then by rank), so every rank computes the same plan.
    """
    order = sorted(range(len(costs)), key=lambda i: (-float(costs[i]), i))
    load = [0.0] * world
    plan: list[list[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        plan[r].append(i)
        load[r] += float(costs[i])
    return plan


def rows_for_rank(costs: list[float], world: int, rank: int) -> list[int]:
    """The share of one rank (what `predict._rank_shard` workers take; also used by the gloo test)."""
    return assign_rows(costs, world)[rank]


def recording_costs(paths) -> list[float]:
    """Cost proxy = file size in bytes (PCM payload ~ duration); missing files cost 0 and fail later, per row."""
    out = []
    for p in paths:
        try:
            out.append(float(os.path.getsize(p)))
        except OSError:
            out.append(0.0)
    return out


def gather_to_rank0(obj, world: int, rank: int):
    """Host-side gather of python objects (torch.distributed, any backend). Returns the list on rank 0, None elsewhere.

    The product's table run needs no gather at all (every worker writes its own label files); this helper exists for callers that
    want the per-recording results in one place and is exercised by tests/test_dist_cpu.py only."""
    if world == 1:
        return [obj]
    import torch.distributed as dist

    bucket = [None] * world if rank == 0 else None
    dist.gather_object(obj, bucket, dst=0)
    return bucket
