"""``fireflies/postprocessing/gauss_blur.py`` -- kornia ``gaussian_blur2d`` semantics (reflect border, separable)
evaluated by the TMA-staged blur kernel."""
import numpy as np

from . import base


class GaussianBlur(base.BasePostProcessingFunction):
    def __init__(self, kernel_size, sigma, probability: float):
        super().__init__(probability)
        self._kernel_size = kernel_size
        self._sigma = sigma

    def spec(self):
        return (tuple(int(k) for k in self._kernel_size), tuple(float(s) for s in self._sigma))

    def post_process(self, image: np.ndarray) -> np.ndarray:
        x = self._to_device(image)
        return base.run_postprocess(x, blur=self.spec())[0].cpu().numpy()
