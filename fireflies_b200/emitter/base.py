"""``fireflies/emitter/base.py``."""
import torch

from .. import entity


class Light(entity.Transformable):
    def __init__(self, name: str, device: torch.device = torch.device("cuda")):
        super().__init__(name, device)
