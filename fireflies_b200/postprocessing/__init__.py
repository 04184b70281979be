from .base import BasePostProcessingFunction
from .white_noise import WhiteNoise
from .postprocessor import PostProcessor
from .gauss_blur import GaussianBlur
from .apply_silhouette import ApplySilhouette
