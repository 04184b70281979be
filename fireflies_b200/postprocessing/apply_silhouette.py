"""``fireflies/postprocessing/apply_silhouette.py`` -- filled circle (cv2) -> 11x11 sigma-5 blur -> multiply.
'Next' row of SURVEY.md section 8(f): the blur runs in the B200 kernel, the circle mask is still drawn by
cv2 on the host exactly like the reference."""
import random

import numpy as np
import torch

from . import base


class ApplySilhouette(base.BasePostProcessingFunction):
    def __init__(self, probability: float = 2.0):
        super().__init__(probability)

    def post_process(self, image: np.ndarray) -> np.ndarray:
        import cv2
        silhouette = np.zeros_like(image)
        cc_x = random.randint(100, 200)
        cc_y = random.randint(200, 300)
        radius = random.randint(170, 230)
        silhouette = cv2.circle(silhouette, (cc_x, cc_y), radius, color=1, thickness=-1)
        x = self._to_device(silhouette)
        blurred = base.run_postprocess(x, blur=((11, 11), (5.0, 5.0)))[0]
        return (torch.from_numpy(image).to(blurred.device) * blurred).cpu().numpy()
