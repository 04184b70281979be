"""Data-parallel plumbing for the scene-sample axis.

Scene samples are independent, so the path shards with NO data-path collective: global sample ``i`` of a step
goes to rank ``i // ceil(total / world)``; the laser pattern and the entity tables are replicated.  The only
exchange is one allreduce(sum) of ``d loss / d points`` ([N,2] fp32, 32 KiB at N=4096) per optimisation step.
Randomisation is keyed by the *global* sample index, so results are bit-identical for any world size.
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def shard_samples(total: int, rank: int, world: int) -> Tuple[int, int]:
    """(first global sample index, count) owned by ``rank`` when ``total`` samples are split over ``world`` ranks
    in contiguous blocks; trailing ranks may own fewer (or zero) samples."""
    if total < 0 or world <= 0 or not (0 <= rank < world):
        raise ValueError("shard_samples: bad arguments")
    per = -(-total // world)
    first = min(rank * per, total)
    return first, min(per, total - first)


def allreduce_sum_(t: torch.Tensor, group=None) -> torch.Tensor:
    """In-place sum over ranks (NCCL on GPUs, gloo in the CPU tests); a no-op without a process group."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def max_over_ranks(value: float, device, group=None) -> float:
    """Timing helper: max of a host scalar over ranks."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        t = torch.tensor([value], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
        return float(t.item())
    return float(value)
