"""Multi-GPU sharding of the MSM path: one process per GPU, the (scalar, base)
pairs split by INDEX RANGE, each rank reduces its slice to one point, and the partial points are
exchanged with a single all-gather (NCCL has no group-law reduction, so all-reduce is realised as
all-gather + local add of `world` points).  The key slice a rank owns is fixed for the life of the
key, so only scalars move.

The functions here are backend-agnostic (`nccl` on GPUs, `gloo` in the CPU tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous index range [lo, hi) of rank `rank`; ranges tile [0, n) and differ by <= 1."""
    lo = (n * rank) // world
    hi = (n * (rank + 1)) // world
    return lo, hi


def all_gather_partials(partial: torch.Tensor, group=None) -> torch.Tensor:
    """partial: uint8[96] Jacobian point of this rank -> uint8[world*96] on every rank."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return partial.clone()
    out = torch.empty(96 * world, dtype=torch.uint8, device=partial.device)
    dist.all_gather_into_tensor(out, partial.contiguous(), group=group)
    return out
