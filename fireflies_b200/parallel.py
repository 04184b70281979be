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


class FoldAllreduce:
    """``sum over ranks of x.sum(0)`` for per-sample pattern gradients ``x [B, ...]`` in ONE kernel over NVLink peer memory
    (``ffb_fold_allreduce``): the fold over this rank's samples, a push of the partial sums into every peer's symmetric
    receive buffer, per-CTA flags and the final sum in rank order.  Replaces ``reduce_over_samples`` + NCCL allreduce for the
    path's only exchange (32 KiB at N = 4096; latency-bound).  Needs one process per GPU of one NVLink domain (<= 8 ranks)
    and torch's symmetric memory; :func:`fold_allreduce` falls back to the NCCL form otherwise."""

    def __init__(self, row_elems: int, device, group=None):
        import ctypes as C
        import torch.distributed._symmetric_memory as symm_mem
        from . import _native as nat
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        if self.world > 8:
            raise RuntimeError("FoldAllreduce: at most 8 ranks (one NVLink domain)")
        self.row = int(row_elems)
        self.ncta = (self.row + 31) // 32
        self.recv = symm_mem.empty(2 * self.world * self.row, dtype=torch.float32, device=device)
        self.flags = symm_mem.empty(self.world * self.ncta, dtype=torch.int32, device=device)
        self.flags.zero_()
        h_recv = symm_mem.rendezvous(self.recv, self.group)
        h_flag = symm_mem.rendezvous(self.flags, self.group)
        self._recv_ptrs = (C.c_void_p * self.world)(*[int(p) for p in h_recv.buffer_ptrs])
        self._flag_ptrs = (C.c_void_p * self.world)(*[int(p) for p in h_flag.buffer_ptrs])
        self._handles = (h_recv, h_flag)
        self.err = torch.zeros(1, dtype=torch.int32, device=device)
        self.epoch = 0
        self._nat, self._C = nat, C
        torch.cuda.synchronize(device)
        dist.barrier(self.group)                 # every rank's flags are zero before anybody's first epoch

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        nat, C = self._nat, self._C
        x = nat.require_cuda(x, torch.float32, "x")
        if x[0].numel() != self.row:
            raise ValueError("FoldAllreduce: row size differs from the one it was built for")
        self.epoch += 1
        out = torch.empty(x.shape[1:], dtype=torch.float32, device=x.device)
        nat.check(nat.lib().ffb_fold_allreduce(x.data_ptr(), x.shape[0], self.row, C.cast(self._recv_ptrs, C.c_void_p),
                                               C.cast(self._flag_ptrs, C.c_void_p), self.rank, self.world, self.epoch,
                                               out.data_ptr(), self.err.data_ptr(), nat.stream()), "ffb_fold_allreduce")
        nat.count()
        return out

    def check(self) -> None:
        """Host sync: raises if a peer ever failed to arrive within the kernel's spin bound."""
        if int(self.err.item()) != 0:
            raise RuntimeError("ffb_fold_allreduce: a peer did not arrive (spin bound exceeded)")


_FOLDERS = {}


def fold_allreduce(x: torch.Tensor, group=None) -> torch.Tensor:
    """``x.sum(0)`` over this rank's samples, summed over all ranks.  One rank: the fold kernel.  Several ranks on NCCL:
    the fused peer-memory kernel (:class:`FoldAllreduce`) unless ``FFB_SYMM_ALLREDUCE=0`` or symmetric memory is unavailable,
    in which case fold + ``dist.all_reduce``."""
    import os
    from .graphics import rasterization as R
    multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    if multi and x.is_cuda and os.environ.get("FFB_SYMM_ALLREDUCE", "1") != "0" and dist.get_backend(group) == "nccl":
        key = (x[0].numel(), x.device.index, id(group))
        f = _FOLDERS.get(key)
        if f is None and key not in _FOLDERS:
            try:
                f = FoldAllreduce(x[0].numel(), x.device, group) if x[0].numel() <= 37888 and dist.get_world_size(group) <= 8 else None
            except Exception as exc:  # noqa: BLE001  (no symmetric memory on this system: keep the NCCL form)
                import warnings
                warnings.warn(f"fireflies_b200: peer-memory allreduce unavailable ({exc}); using NCCL")
                f = None
            _FOLDERS[key] = f
        if f is not None:
            return f(x)
    out = R.reduce_over_samples(x) if x.shape[0] > 1 else x[0].clone()
    return allreduce_sum_(out, group)
