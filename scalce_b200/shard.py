"""Host-side sharding helpers for the one-process-per-GPU launch (torchrun).

Reads shard across ranks as contiguous input slices: rank g owns
[bounds[g], bounds[g+1]); mates travel with mate 1. Timing is reported as the maximum over
ranks, throughput as total units / that maximum.
"""
from __future__ import annotations


def shard_bounds(n: int, world: int):
    """Contiguous, balanced slices: the first n % world ranks get one extra read."""
    base, extra = divmod(n, world)
    b = [0]
    for g in range(world):
        b.append(b[-1] + base + (1 if g < extra else 0))
    return b


def max_over_ranks(values, dist=None, device=None):
    """Element-wise MAX of a list of floats over all ranks (identity when not distributed)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return list(values)
    import torch
    t = torch.tensor(list(values), dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t]


def sum_over_ranks(values, dist=None, device=None):
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return list(values)
    import torch
    t = torch.tensor(list(values), dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [float(x) for x in t]
