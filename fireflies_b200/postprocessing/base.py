"""``fireflies/postprocessing/base.py`` + the bridge to ``ffb_postprocess``."""
from __future__ import annotations

import ctypes as C
import random
from typing import Optional

import numpy as np
import torch

from .. import _native as nat


def run_postprocess(frames: torch.Tensor, blur=None, noise=None, gates: Optional[torch.Tensor] = None,
                    noise_injected: Optional[torch.Tensor] = None, seed: int = 0, frame0: int = 0) -> torch.Tensor:
    """``frames`` f32 CUDA ``[B,H,W]`` -> new tensor.  ``blur`` = ((ky,kx),(sy,sx)) or None; ``noise`` = (mean,std)
    or None; ``gates`` u8 ``[B,2]`` (blur gate, noise gate) or None = all on; ``noise_injected`` f64 ``[B,H,W]``."""
    frames = nat.require_cuda(frames, torch.float32, "frames")
    B, H, W = frames.shape
    d = nat.PostDesc(B, H, W, 0, 0, 0.0, 0.0, 0, 0.0, 0.0, seed & (2**64 - 1), frame0)
    if blur is not None:
        (d.blur_ky, d.blur_kx), (d.blur_sy, d.blur_sx) = blur
    if noise is not None:
        d.noise, d.noise_mean, d.noise_std = 1, float(noise[0]), float(noise[1])
    if gates is not None:
        gates = nat.require_cuda(gates, torch.uint8, "gates")
    if noise_injected is not None:
        noise_injected = nat.require_cuda(noise_injected, torch.float64, "noise_injected")
    out = torch.empty_like(frames)
    nat.check(nat.lib().ffb_postprocess(C.byref(d), frames.data_ptr(), nat.ptr(gates), nat.ptr(noise_injected),
                                        out.data_ptr(), nat.stream()), "ffb_postprocess")
    nat.count()
    return out


class BasePostProcessingFunction:
    def __init__(self, probability: float):
        self._probability = probability

    def apply(self, image: np.ndarray) -> np.ndarray:
        if random.uniform(0, 1) < self._probability:          # base.py:10-14
            return self.post_process(image)
        return image

    def gate(self) -> bool:
        """One Bernoulli draw from python's ``random`` -- the same stream ``apply`` consumes."""
        return random.uniform(0, 1) < self._probability

    def post_process(self, image: np.ndarray) -> np.ndarray:
        raise NotImplementedError

    @staticmethod
    def _to_device(image: np.ndarray, device="cuda") -> torch.Tensor:
        return torch.from_numpy(np.ascontiguousarray(image, dtype=np.float32)).to(device, non_blocking=True).unsqueeze(0)
