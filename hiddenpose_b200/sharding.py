"""Data-parallel plumbing for the LCT path: one process per GPU, sharded by transient.

Every (batch, channel) volume is an independent linear transform with shared read-only
constants (/root/reference/models/tflct.py:121 flattens B*D and nothing mixes channels),
so ranks split the batch contiguously and the data path needs no collective.  The only
collective is the timing reduction (max over ranks), plus whatever the caller's training
loop does with the gradients of its own parameters (the layer has none).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(total: int, world: int, rank: int):
    """Contiguous split of `total` transients over `world` ranks; the first `total % world`
    ranks take one extra.  Returns (begin, end)."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    base, extra = divmod(int(total), int(world))
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def shard_batch(x: torch.Tensor, world: int, rank: int) -> torch.Tensor:
    b, e = shard_bounds(x.shape[0], world, rank)
    return x[b:e]


def max_over_ranks(value: float, device=None) -> float:
    """Max of a per-rank scalar (the timing rule for multi-GPU numbers)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_batch(y_local: torch.Tensor, total: int):
    """All-gather per-rank output shards (possibly ragged) back into batch order on every rank."""
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = [shard_bounds(total, world, r) for r in range(world)]
    longest = max(e - b for b, e in sizes)
    pad = torch.zeros((longest,) + tuple(y_local.shape[1:]), dtype=y_local.dtype, device=y_local.device)
    pad[: y_local.shape[0]] = y_local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad)
    return torch.cat([p[: e - b] for p, (b, e) in zip(parts, sizes)], dim=0)
