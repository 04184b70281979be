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
        if x.shape[0] == 0:                     # a rank without samples still takes part in the exchange: it contributes zeros
            x = torch.zeros((1,) + tuple(x.shape[1:]), dtype=torch.float32, device=x.device)
        if x[0].numel() != self.row:
            raise ValueError("FoldAllreduce: row size differs from the one it was built for")
        if x.device != self.recv.device:
            raise ValueError("FoldAllreduce: tensor lives on another device than the receive buffers")
        self.epoch += 1
        with torch.cuda.device(x.device):       # the launch uses the current device's stream
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


def _peer_path_available(x: torch.Tensor, group) -> bool:
    """Whether EVERY rank can take the peer-memory path.  The choice must be collective: a rank that fell back to NCCL on its own
    would leave its peers spinning in the flag wait of the kernel."""
    import os
    import importlib
    import math
    ok = x.is_cuda and os.environ.get("FFB_SYMM_ALLREDUCE", "0") == "1" and dist.get_backend(group) == "nccl"
    ok = ok and math.prod(x.shape[1:]) <= 37888 and dist.get_world_size(group) <= 8
    if ok:
        try:
            importlib.import_module("torch.distributed._symmetric_memory")
        except Exception:  # noqa: BLE001
            ok = False
    flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=x.device)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
    return bool(flag.item())


def fold_allreduce(x: torch.Tensor, group=None) -> torch.Tensor:
    """``x.sum(0)`` over this rank's samples, summed over all ranks.  One rank: the fold kernel.  Several ranks: fold +
    ``dist.all_reduce`` (NCCL over NVLink; 32 KiB, latency-bound).  ``FFB_SYMM_ALLREDUCE=1`` switches to the fused peer-memory
    kernel (:class:`FoldAllreduce`) when every rank can take it (agreed once per tensor shape with an allreduce(MIN)).  Measured on
    2 x B200 inside the step (profiles/r02m): 3.95 ms per step with NCCL against 4.95 - 6.6 ms with the peer-memory kernel, whose
    flag spin interacts badly with the rest of the step although it is twice as fast in isolation (28 against 52 us) -- NCCL is the
    default, the kernel stays opt-in and covered by tests/test_multi_gpu.py.  A rank with zero samples contributes zeros."""
    from .graphics import rasterization as R
    multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    if multi and x.is_cuda:
        import math
        row = math.prod(x.shape[1:])
        key = (row, x.device.index, id(group))
        if key not in _FOLDERS:
            # construction is collective (symmetric-memory rendezvous + barrier): an error here is raised, not papered over
            _FOLDERS[key] = FoldAllreduce(row, x.device, group) if _peer_path_available(x, group) else None
        f = _FOLDERS[key]
        if f is not None:
            return f(x)
    if x.shape[0] == 0:
        out = torch.zeros(x.shape[1:], dtype=torch.float32, device=x.device)
    else:
        out = R.reduce_over_samples(x) if x.shape[0] > 1 else x[0].clone()
    return allreduce_sum_(out, group)


def check_folders() -> None:
    """Raises if any peer-memory exchange so far missed a peer (host sync; :meth:`FoldAllreduce.check`)."""
    for f in _FOLDERS.values():
        if f is not None:
            f.check()


def multi_gpu_selfcheck(device, group=None, n_points: int = 512, texture=(256, 256), sigma: float = 25.0, per_rank: int = 4,
                        mesh_vertices: int = 3000, rounds: int = 6) -> dict:
    """Correctness of the N > 1 path on real devices; every rank calls it.  Small shapes, a few milliseconds.

    1. ``allreduce``: :func:`fold_allreduce` of this rank's per-sample pattern gradients equals fold + ``dist.all_reduce`` (1e-5
       relative) and is bit-identical on every rank (all_gather), ``rounds`` times in a row (epoch double buffering), the kernel's
       missing-peer flag checked every round.
    2. ``rng_split_invariant``: the scene samples of rank 0's shard recomputed on this rank are bit-identical to rank 0's own
       (broadcast), and this rank's shard equals the same rows of a full-batch randomisation.
    3. ``sharded_gradient``: the allreduced gradient of the sharded step equals the full batch evaluated on this rank alone (1e-4
       relative to the gradient's norm).
    Returns ``{check: "ok"}``; raises ``AssertionError`` naming the first failing check."""
    import fireflies_b200 as ff
    from .graphics import rasterization as R
    multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    world = dist.get_world_size(group) if multi else 1
    rank = dist.get_rank(group) if multi else 0
    device = torch.device(device)
    total = per_rank * world
    first, n = shard_samples(total, rank, world)
    out = {}
    # ---- scene: one mesh with ranged translation / rotation / scale, Philox train mode ----
    class _P(dict):
        def update(self, *a, **k):
            if a or k:
                return super().update(*a, **k)
    verts = torch.rand(mesh_vertices, 3, generator=torch.Generator().manual_seed(21)) * 2 - 1
    sc = ff.Scene(_P(), device=device)
    m = ff.entity.Mesh("mesh-selfcheck", verts.to(device), device)
    c = lambda v: torch.tensor(v, device=device)  # noqa: E731
    m.translate(c([-0.5, -0.5, -0.5]), c([0.5, 0.5, 0.5]))
    m.rotate(c([-3.1, -3.1, -3.1]), c([3.1, 3.1, 3.1]))
    m.scale(c([0.5, 0.5, 0.5]), c([2.0, 2.0, 2.0]))
    sc._meshes.append(m)
    sc.train()
    sb = sc.batch(seed=4321)
    own = sb.randomize(n, sample0=first)
    shard0 = sb.randomize(per_rank, sample0=0)
    ref = shard0.vertices.clone()
    if multi:
        dist.broadcast(ref, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    full = sb.randomize(total, sample0=0)
    ok = torch.equal(ref, shard0.vertices) and torch.equal(full.vertices[first:first + n], own.vertices) \
        and torch.equal(full.world[first:first + n], own.world)
    assert ok, "rng_split_invariant: randomised scene samples depend on the rank layout"
    out["rng_split_invariant"] = "ok"
    # ---- pattern gradients: per-sample upstream gradients keyed by the GLOBAL sample index ----
    ts0, ts1 = int(texture[0]), int(texture[1])
    pattern = (torch.rand(n_points, 2, generator=torch.Generator().manual_seed(3)) * 0.9 + 0.05).to(device)
    def upstream(i):
        g = torch.Generator().manual_seed(1000 + i)
        return torch.randn(ts0, ts1, generator=g), torch.randn(ts1, ts0, generator=g)
    ups = [upstream(i) for i in range(total)]
    gS = torch.stack([u[0] for u in ups]).to(device)
    gO = torch.stack([u[1] for u in ups]).to(device)
    def per_sample_grads(lo, cnt):
        pts = pattern.unsqueeze(0).repeat(cnt, 1, 1).contiguous()
        plan = R._SplatPlan(pts, cnt, sigma, ts0, ts1, 4, 5)
        plan.forward(pts, True, True, True)
        return plan.backward(pts, gS[lo:lo + cnt].contiguous(), gO[lo:lo + cnt].contiguous(), True)
    d_own = per_sample_grads(first, n)
    gathered = None
    for _ in range(rounds):
        a = fold_allreduce(d_own, group)
        b = R.reduce_over_samples(d_own) if n > 1 else d_own[0].clone()
        allreduce_sum_(b, group)
        scale = float(b.abs().max())
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-5 * scale), "allreduce: peer-memory exchange differs from fold + NCCL allreduce"
        if multi:
            gathered = [torch.empty_like(a) for _ in range(world)]
            dist.all_gather(gathered, a.contiguous(), group=group)
            assert all(torch.equal(gathered[0], g) for g in gathered), "allreduce: result is not bit-identical on every rank"
        check_folders()
    out["allreduce"] = "ok"
    d_full = per_sample_grads(0, total).double().sum(0)
    err = float((a.double() - d_full).norm() / d_full.norm())
    assert err < 1e-4, f"sharded_gradient: allreduced gradient differs from the full batch by {err:.2e} of its norm"
    out["sharded_gradient"] = "ok"
    out["ranks"] = world
    out["exchange"] = "peer-memory kernel" if any(f is not None for f in _FOLDERS.values()) else ("nccl" if multi else "single rank")
    return out
