"""``fireflies/postprocessing/apply_silhouette.py`` -- filled disc -> 11x11 sigma-5 blur -> multiply, same interface.

The disc parameters are drawn exactly like the reference (``random.randint`` x3, apply_silhouette.py:23-25); disc, blur
and product run on the device (``ffb_silhouette``).  Deviation: the disc is the analytic set
``(x-cx)^2 + (y-cy)^2 <= r^2`` instead of ``cv2.circle``'s rasterisation (OpenCV is not a dependency here); the two
differ on boundary texels only, and those are smoothed by the 11x11 blur that follows.
"""
import random
from typing import Optional, Sequence

import numpy as np
import torch

from .. import _native as nat
from . import base


def run_silhouette(frames: torch.Tensor, discs: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``frames`` f32 CUDA ``[B,H,W]``, ``discs`` int32 ``[B,3]`` = (cx, cy, r) -> ``frames * blur(disc)``."""
    frames = nat.require_cuda(frames, torch.float32, "frames")
    discs = nat.require_cuda(discs, torch.int32, "discs")
    B, H, W = frames.shape
    if discs.shape != (B, 3):
        raise ValueError("discs must be [B, 3]")
    out = torch.empty_like(frames) if out is None else out
    mask = torch.empty_like(frames)
    nat.check(nat.lib().ffb_silhouette(frames.data_ptr(), discs.data_ptr(), B, H, W, mask.data_ptr(), out.data_ptr(), nat.stream()),
              "ffb_silhouette")
    nat.count(2)
    return out


class ApplySilhouette(base.BasePostProcessingFunction):
    def __init__(self, probability: float = 2.0):
        super().__init__(probability)

    @staticmethod
    def draw_disc() -> Sequence[int]:
        cc_x = random.randint(100, 200)
        cc_y = random.randint(200, 300)
        radius = random.randint(170, 230)
        return cc_x, cc_y, radius

    def post_process(self, image: np.ndarray) -> np.ndarray:
        x = self._to_device(image)
        discs = torch.tensor([self.draw_disc()], dtype=torch.int32, device=x.device)
        return run_silhouette(x, discs)[0].cpu().numpy()
