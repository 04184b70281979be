"""``fireflies/postprocessing/postprocessor.py`` + the batched device path."""
from __future__ import annotations

from typing import List, Optional

import numpy as np
import torch

from . import base
from .gauss_blur import GaussianBlur
from .white_noise import WhiteNoise


class PostProcessor:
    def __init__(self, post_process_funcs: List[base.BasePostProcessingFunction]):
        self._post_process_functs = post_process_funcs

    def post_process(self, image: np.ndarray) -> np.ndarray:
        """postprocessor.py:14-19: copy, then ``apply`` each function in order (numpy in, numpy out)."""
        image_copy = image.copy()
        for func in self._post_process_functs:
            image_copy = func.apply(image_copy)
        return image_copy

    def _fusable(self):
        fs = self._post_process_functs
        blur = [f for f in fs if isinstance(f, GaussianBlur)]
        noise = [f for f in fs if isinstance(f, WhiteNoise)]
        ok = len(blur) <= 1 and len(noise) <= 1 and len(blur) + len(noise) == len(fs)
        if ok and blur and noise:
            ok = fs.index(blur[0]) < fs.index(noise[0])
        return ok, (blur[0] if blur else None), (noise[0] if noise else None)

    def post_process_batch(self, frames: torch.Tensor, seed: int = 0, frame0: int = 0,
                           gates: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Device-resident batch ``[B,H,W]`` through the chain ``[GaussianBlur?, WhiteNoise?]`` in ONE fused
        launch.  Gates: drawn here from python's ``random`` in the reference's order (frame-major, one draw
        per function, base.py:10-14) unless given as u8 ``[B,2]``.  Noise comes from the in-kernel Philox
        stream keyed by (seed, frame0 + b, texel)."""
        ok, blur, noise = self._fusable()
        if not ok:
            raise NotImplementedError("post_process_batch fuses the chain [GaussianBlur?, WhiteNoise?] only")
        B = frames.shape[0]
        if gates is None:
            g = np.ones((B, 2), dtype=np.uint8)
            for b in range(B):
                for f in self._post_process_functs:
                    g[b, 0 if isinstance(f, GaussianBlur) else 1] = f.gate()
            gates = torch.from_numpy(g).to(frames.device)
        return base.run_postprocess(frames, blur=blur.spec() if blur else None, noise=noise.spec() if noise else None,
                                    gates=gates, seed=seed, frame0=frame0)
