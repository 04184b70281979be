"""``fireflies/sampling/uniform_integer.py``.  The reference constructor passes the *builtins* ``min``/``max``
to the base class and raises (uniform_integer.py:17); this implements the documented
intent: integers from ``[min_integer, max_integer)`` like ``range()``."""
import random

import torch


class UniformIntegerSampler:
    def __init__(self, min_integer: int, max_integer: int, eval_step_size: int = 1,
                 device: torch.device = torch.device("cuda")) -> None:
        self._device = device
        self._train = True
        self._min_range, self._max_range = int(min_integer), int(max_integer)
        self._eval_step_size = eval_step_size
        self._current_step = 0

    def train(self) -> None:
        self._train = True

    def eval(self) -> None:
        self._train = False

    def sample(self) -> int:
        return self.sample_train() if self._train else self.sample_eval()

    def sample_eval(self) -> int:                      # uniform_integer.py:20-27
        sample = self._current_step
        self._current_step += self._eval_step_size
        if self._current_step >= self._max_range:
            self._current_step = self._min_range
        return sample

    def sample_train(self) -> int:                     # uniform_integer.py:29-30
        return random.randint(0, self._max_range - 1)
