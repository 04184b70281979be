"""``fireflies/entity/curve.py`` -- NURBS-curve camera paths, same interface.

The reference's constructor raises (``super(Curve, self).__init__(self, name, device)`` passes ``self`` twice,
entity/curve.py:24) and its evaluator lives in the un-vendored geomdl 5.3.1; the evident intent is implemented: a
Transformable whose ``randomize()`` walks a NURBS curve.  Point evaluation (fp64, like geomdl's Python floats), the
Rodrigues rotation from the path tangent and the ``T @ R @ W`` product are one launch (``ffb_curve_pose``).
Kept quirks: ``train()`` draws ``random.uniform(0 + curve_epsilon, eval_interval_start)`` -- both 0.05, so training
poses sit at the start of the path (:79-82); ``eval()`` resets ``_curve_delta``, an attribute nothing reads (:44), so the
eval walk continues from wherever ``curve_delta`` stands; no ``randomizable()`` check (:78).
"""
from __future__ import annotations

import random

import torch

from ..utils.nurbs import NurbsCurve, as_nurbs
from . import base


class Curve(base.Transformable):
    count = 0.0

    def fromObj(path):
        """Unimplemented in the reference as well (``pass``, entity/curve.py:14-16); see ``utils.io.importBlenderNurbsObj``."""
        pass

    def __init__(self, name: str, curve, device: torch.device = torch.device("cuda")):
        super().__init__(name, device)
        self._curve: NurbsCurve = as_nurbs(curve, device)
        self.curve_epsilon = 0.05
        self.curve_delta = self.curve_epsilon
        self._interp_steps = 1000
        self._interp_delta = 1.0 / self._interp_steps
        self.eval_interval_start = 0.05

    def train(self) -> None:
        self._train = True
        self._continuous = False

    def eval(self) -> None:
        self._train = False
        self._continuous = True
        self._curve_delta = self.eval_interval_start

    def setContinuous(self, continuous: bool) -> None:
        self._continuous = continuous

    def sample_rotation(self) -> torch.Tensor:
        """entity/curve.py:48-69."""
        return self._curve.poses([self.curve_delta], self._world, parts=True)[1][0]

    def sample_translation(self) -> torch.Tensor:
        """entity/curve.py:71-80."""
        return self._curve.poses([self.curve_delta], self._world, parts=True)[2][0]

    def _advance(self) -> float:
        if self._train:
            self.curve_delta = random.uniform(0 + self.curve_epsilon, self.eval_interval_start)
        else:
            self.curve_delta += self._interp_delta
            if self.curve_delta > 1.0 - self.curve_epsilon:
                self.curve_delta = self.eval_interval_start
        return self.curve_delta

    def randomize(self) -> None:
        """entity/curve.py:82-96."""
        self._randomized_world = self._curve.poses([self._advance()], self._world)[0]

    def randomize_batch(self, count: int) -> torch.Tensor:
        """``count`` consecutive ``randomize()`` steps in one launch -> ``[count,4,4]`` (the last one becomes the entity's
        randomized world).  The path parameters are the ones ``count`` calls of ``randomize()`` would visit."""
        ts = [self._advance() for _ in range(int(count))]
        worlds = self._curve.poses(ts, self._world)
        if count:
            self._randomized_world = worlds[-1]
        return worlds
