"""``fireflies/sampling/uniform.py``."""
import torch

from . import base
from ..utils import math as ffmath


class UniformSampler(base.Sampler):
    def __init__(self, min, max, eval_step_size: float = 0.01, device: torch.device = torch.device("cuda")) -> None:
        super().__init__(min, max, eval_step_size, device)

    def sample_train(self) -> torch.Tensor:
        # uniform.py:16-19 -> randomBetweenTensors: torch.rand from the global generator, affine map in-kernel
        return ffmath.randomBetweenTensors(self._min_range, self._max_range)
