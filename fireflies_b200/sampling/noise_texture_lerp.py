"""``fireflies/sampling/noise_texture_lerp.py`` -- Perlin-noise material textures, same interface.

The octave sum, the min/max normalisation and the colour lerp run in libffb200 (``ffb_perlin_texture``).  The random
draws are made exactly where the reference makes them -- ``random.randint`` / ``random.uniform`` for the lattice
resolution, octave count and persistence, ``torch.rand`` on the CPU generator for the lattice angles
(noise_texture_lerp.py:21,77-79) -- so a seeded run consumes both streams identically.
"""
from __future__ import annotations

import random
from typing import List

import torch

from .. import _native as nat
from . import base


def perlin_angles(res, octaves: int) -> List[torch.Tensor]:
    """The ``torch.rand(res[0]+1, res[1]+1)`` draws of rand_perlin_2d_octaves, octave by octave (:21, :57-61)."""
    out, f = [], 1
    for _ in range(octaves):
        out.append(torch.rand(f * res[0] + 1, f * res[1] + 1))
        f *= 2
    return out


def perlin_texture(shape, res, octaves: int, persistence: float, angles: List[torch.Tensor], color_a=None, color_b=None,
                   device=torch.device("cuda")):
    """``rand_perlin_2d_octaves(shape, res, octaves, persistence)`` from given lattice draws -> ``(noise [H,W], texture
    [3,H,W] or None)``; the texture is ``lerp(color_a, color_b, (noise - min) / (max - min))`` (:80-98)."""
    H, W = int(shape[0]), int(shape[1])
    ang = torch.cat([a.reshape(-1).float() for a in angles]).to(device)
    noise = torch.empty((H, W), dtype=torch.float32, device=device)
    mm = torch.empty(2, dtype=torch.int32, device=device)
    out = ca = cb = None
    if color_a is not None:
        ca = nat.require_cuda(color_a.detach().float().reshape(3).contiguous().to(device), torch.float32, "color_a")
        cb = nat.require_cuda(color_b.detach().float().reshape(3).contiguous().to(device), torch.float32, "color_b")
        out = torch.empty((3, H, W), dtype=torch.float32, device=device)
    nat.check(nat.lib().ffb_perlin_texture(ang.data_ptr(), H, W, int(res[0]), int(res[1]), int(octaves), float(persistence),
                                           nat.ptr(ca), nat.ptr(cb), noise.data_ptr(), mm.data_ptr(), nat.ptr(out), nat.stream()),
              "ffb_perlin_texture")
    nat.count(2 if out is not None else 1)
    return noise, out


def rand_perlin_2d(shape, res, fade=None, device=torch.device("cuda")) -> torch.Tensor:
    """noise_texture_lerp.py:8-50: one octave; draws ``torch.rand(res[0]+1, res[1]+1)`` like the reference.  Only the
    reference's default quintic fade is compiled into the kernel."""
    if fade is not None:
        raise NotImplementedError("rand_perlin_2d: only the default fade 6t^5 - 15t^4 + 10t^3 is supported")
    return perlin_texture(shape, res, 1, 1.0, perlin_angles(res, 1), device=device)[0]


def rand_perlin_2d_octaves(shape, res, octaves=1, persistence=0.5, device=torch.device("cuda")) -> torch.Tensor:
    """noise_texture_lerp.py:53-62."""
    return perlin_texture(shape, res, octaves, persistence, perlin_angles(res, octaves), device=device)[0]


class NoiseTextureLerpSampler(base.Sampler):
    _KIND = None          # no device record: its samples are textures, not the 1-3 scalars the batched sampler table holds
    def __init__(self, color_a: torch.Tensor, color_b: torch.Tensor, texture_shape: List[int], eval_step_size: float = 0.01,
                 device: torch.device = torch.device("cuda")) -> None:
        super().__init__(torch.tensor([0.0], device=device), torch.tensor([1.0], device=device), eval_step_size, device)
        self._color_a = color_a
        self._color_b = color_b
        self._texture_shape = texture_shape

    def sample_train(self) -> torch.Tensor:
        i = 2 ** random.randint(1, 6)
        octaves = random.randint(1, 4)
        persistence = random.uniform(0.1, 2.0)
        angles = perlin_angles((i, i), octaves)
        return perlin_texture(self._texture_shape, (i, i), octaves, persistence, angles, self._color_a, self._color_b, self._device)[1]

    # the reference's eval mode is its train mode (:100-102)
    def sample_eval(self) -> torch.Tensor:
        return self.sample_train()
