"""``fireflies/sampling/base.py`` -- sampler base class, state kept in one device-resident
``ffb_sampler`` record (include/ffb200.h) so the batched kernels can read it without a host round trip.

``_min_range`` / ``_max_range`` / ``_current_step`` are *views into that record*: the in-place edits the
reference makes through ``get_min()[i] = v`` (entity/base.py:141-146) land directly in kernel-visible memory.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from .. import _native as nat

_WORDS = C.sizeof(nat.Sampler) // 4          # 19 x 4-byte fields
_OFF_MIN, _OFF_MAX, _OFF_CUR, _OFF_MEAN, _OFF_STD = 4, 7, 10, 13, 16


def _lerp_native(a: torch.Tensor, b: torch.Tensor, u: torch.Tensor) -> torch.Tensor:
    """``u*(b-a)+a`` in libffb200 (utils/math.py:174-175)."""
    a_ = nat.require_cuda(a.detach().float().contiguous(), torch.float32, "a")
    b_ = nat.require_cuda(b.detach().float().contiguous(), torch.float32, "b")
    u_ = nat.require_cuda(u.float().contiguous(), torch.float32, "u")
    out = torch.empty_like(a_)
    nat.check(nat.lib().ffb_uniform_between(a_.data_ptr(), b_.data_ptr(), u_.data_ptr(), a_.numel(), out.data_ptr(),
                                            nat.stream()), "ffb_uniform_between")
    nat.count()
    return out


class Sampler:
    _KIND = nat.SAMPLER_UNIFORM

    def __init__(self, min, max, eval_step_size: float = 0.01, device: torch.device = torch.device("cuda")) -> None:
        self._device = device
        mn = min.detach().clone().float().reshape(-1) if type(min) is torch.Tensor else torch.tensor([min], dtype=torch.float32)
        mx = max.detach().clone().float().reshape(-1) if type(max) is torch.Tensor else torch.tensor([max], dtype=torch.float32)
        if mn.numel() != mx.numel() or mn.numel() not in (1, 3):
            raise ValueError("fireflies_b200 samplers hold 1 or 3 components")
        dim = mn.numel()
        dev = mn.device if (type(min) is torch.Tensor and mn.is_cuda) else torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("fireflies_b200 samplers live on a CUDA device; there is no CPU path")
        rec = np.zeros(_WORDS, dtype=np.int32)
        rec[0], rec[1], rec[2] = (nat.SAMPLER_UNIFORM if self._KIND is None else self._KIND), dim, 0
        rec[3:4] = np.array([eval_step_size], dtype=np.float32).view(np.int32)
        self._buf = torch.from_numpy(rec).to(dev)
        fv = self._buf.view(torch.float32)
        self._min_range = fv[_OFF_MIN:_OFF_MIN + dim]
        self._max_range = fv[_OFF_MAX:_OFF_MAX + dim]
        self._current_step = fv[_OFF_CUR:_OFF_CUR + dim]       # cloned from min at construction (base.py:26-30)
        self._min_range.copy_(mn)
        self._max_range.copy_(mx)
        self._current_step.copy_(mn)
        self._dim = dim
        self._train = True
        self._eval_step_size = eval_step_size

    # -- record access for the batched path ---------------------------------------------------------
    def record(self) -> torch.Tensor:
        """int32 [19] device view of this sampler's ``ffb_sampler`` record."""
        return self._buf

    def set_sample_interval(self, min: torch.Tensor, max: torch.Tensor) -> None:
        self._min_range.copy_(min.reshape(-1))
        self._max_range.copy_(max.reshape(-1))

    def get_min(self) -> torch.Tensor:
        return self._min_range

    def get_max(self) -> torch.Tensor:
        return self._max_range

    def set_sample_max(self, max: torch.Tensor) -> None:
        self._max_range.copy_(max.reshape(-1))

    def set_sample_min(self, min: torch.Tensor) -> None:
        self._min_range.copy_(min.reshape(-1))

    def train(self) -> None:
        self._train = True

    def eval(self) -> None:
        self._train = False

    def sample(self) -> torch.Tensor:
        return self.sample_train() if self._train else self.sample_eval()

    def sample_train(self) -> torch.Tensor:
        raise NotImplementedError

    def _native_sample(self, mode: int, variates=None) -> torch.Tensor:
        out = torch.empty((1, 1, 3), dtype=torch.float32, device=self._buf.device)
        nat.check(nat.lib().ffb_sample(self._buf.data_ptr(), 1, 1, mode, 0, 0, nat.ptr(variates), out.data_ptr(), nat.stream()),
                  "ffb_sample")
        nat.count()
        return out[0, 0]

    def sample_eval(self) -> torch.Tensor:
        """sampling/base.py:64-74, stepped on the device *with* the reference's aliasing (post-increment
        value returned; after the first wrap the range minimum drifts).  Returns a fresh tensor holding the
        value the reference's (aliased) return tensor has at return time."""
        return self._native_sample(nat.MODE_EVAL)[: self._dim].clone()


def rehome(sampler: Sampler, row: torch.Tensor) -> None:
    """Move a sampler's ``ffb_sampler`` record into ``row`` (an int32 [19] slice of a batched sampler table) so the
    batched kernels and the per-object API share one copy of the state."""
    row.copy_(sampler._buf)
    sampler._buf = row
    fv = row.view(torch.float32)
    d = sampler._dim
    sampler._min_range = fv[_OFF_MIN:_OFF_MIN + d]
    sampler._max_range = fv[_OFF_MAX:_OFF_MAX + d]
    sampler._current_step = fv[_OFF_CUR:_OFF_CUR + d]
