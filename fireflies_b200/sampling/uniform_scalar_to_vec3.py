"""``fireflies/sampling/uniform_scalar_to_vec3.py``: one scalar draw replicated three times."""
import torch

from . import base
from .. import _native as nat
from ..utils import math as ffmath


class UniformScalarToVec3Sampler(base.Sampler):
    _KIND = nat.SAMPLER_SCALAR_TO_VEC3

    def __init__(self, min, max, eval_step_size: float = 0.01, device: torch.device = torch.device("cuda")) -> None:
        super().__init__(min, max, eval_step_size, device)

    def sample_train(self) -> torch.Tensor:
        s = ffmath.randomBetweenTensors(self._min_range, self._max_range)      # uniform_scalar_to_vec3.py:18-24
        return s.expand(3).contiguous()

    def sample_eval(self) -> torch.Tensor:
        return self._native_sample(nat.MODE_EVAL).clone()                      # uniform_scalar_to_vec3.py:26-38
