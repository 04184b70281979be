"""``fireflies/postprocessing/white_noise.py``.

``post_process`` keeps the reference's stream: the normal variates come from ``np.random.normal`` (fp64) and are
added on the device in fp64 exactly like numpy does (white_noise.py:17), so results are bit-identical for the same
numpy seed.  The batched path (``PostProcessor.post_process_batch``) draws the noise in-kernel with Philox."""
import numpy as np
import torch

from . import base


class WhiteNoise(base.BasePostProcessingFunction):
    def __init__(self, mean: float, std: float, probability: float):
        super().__init__(probability)
        self._mean = mean
        self._std = std

    def spec(self):
        return (float(self._mean), float(self._std))

    def post_process(self, image: np.ndarray) -> np.ndarray:
        noise = np.random.normal(np.ones_like(image) * self._mean, np.ones_like(image) * self._std)
        x = self._to_device(image)
        n = torch.from_numpy(np.ascontiguousarray(noise, dtype=np.float64)).to(x.device).unsqueeze(0)
        out = base.run_postprocess(x, noise=self.spec(), noise_injected=n)[0].cpu().numpy()
        image[...] = out          # the reference mutates its argument in place (white_noise.py:17)
        return out
