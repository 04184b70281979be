"""``fireflies/sampling/animation.py``: integer frame indices.  Pure-python state like the reference (the
indices are host integers there too); the batched path draws them with ``ffb_sample_anim_index``."""
import random

import torch


class AnimationSampler:
    def __init__(self, min_integer_train: int, max_integer_train: int, min_integer_eval: int, max_integer_eval: int,
                 eval_step_size: int = 1, device: torch.device = torch.device("cuda")) -> None:
        self._device = device
        self._train = True
        self._eval_step_size = eval_step_size
        self._min_integer_train = min_integer_train
        self._max_integer_train = max_integer_train
        self._min_integer_eval = min_integer_eval
        self._max_integer_eval = max_integer_eval
        self._current_step = min_integer_eval

    def train(self) -> None:
        self._train = True

    def eval(self) -> None:
        self._train = False

    def sample(self) -> int:
        return self.sample_train() if self._train else self.sample_eval()

    def sample_eval(self) -> int:                      # animation.py:27-34 (max inclusive)
        sample = self._current_step
        self._current_step += self._eval_step_size
        if self._current_step > self._max_integer_eval:
            self._current_step = self._min_integer_eval
        return sample

    def sample_train(self) -> int:                     # animation.py:36-37
        return random.randint(self._min_integer_train, self._max_integer_train - 1)

    def set_train_interval(self, min_integer_train: int, max_integer_train: int) -> None:
        self._min_integer_train = min_integer_train
        self._max_integer_train = max_integer_train

    def set_eval_interval(self, min_integer_eval: int, max_integer_eval: int) -> None:
        self._min_integer_eval = min_integer_eval
        self._max_integer_eval = max_integer_eval
