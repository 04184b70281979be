"""``fireflies/sampling/gaussian_distribution.py``: train = unclamped normal(mean, std)."""
import torch

from . import base
from .. import _native as nat


class GaussianSampler(base.Sampler):
    _KIND = nat.SAMPLER_GAUSSIAN

    def __init__(self, min, max, mean, std, eval_step_size: float = 0.01, device: torch.device = torch.device("cuda")) -> None:
        super().__init__(min, max, eval_step_size, device)
        self._mean = mean
        self._std = std
        self._sync_moments()

    def _sync_moments(self) -> None:
        fv = self._buf.view(torch.float32)
        m = torch.as_tensor(self._mean, dtype=torch.float32).reshape(-1)
        s = torch.as_tensor(self._std, dtype=torch.float32).reshape(-1)
        fv[base._OFF_MEAN:base._OFF_MEAN + m.numel()].copy_(m)
        fv[base._OFF_STD:base._OFF_STD + s.numel()].copy_(s)

    def sample_train(self) -> torch.Tensor:
        # gaussian_distribution.py:19-20: torch.normal(mean, std) == mean + std * randn (global generator)
        self._sync_moments()
        z = torch.zeros((1, 1, 3), dtype=torch.float32, device=self._buf.device)
        z[0, 0, : self._dim] = torch.randn(self._dim, device=self._buf.device)
        return self._native_sample(nat.MODE_INJECTED, z)[: self._dim].clone()
