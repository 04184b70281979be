"""``fireflies/material/base.py``: attributes only; transform methods warn (and, unlike the reference whose
Translation/World decorators return ``None``, still work -- SURVEY.md section 2 #4)."""
import torch

from .. import entity
from ..utils.warnings import (RotationAssignmentWarning, RelativeAssignmentWarning, TranslationAssignmentWarning,
                              WorldAssignmentWarning)


class Material(entity.Transformable):
    def __init__(self, name: str, device: torch.device = torch.device("cuda")):
        super().__init__(name, device)

    def randomize(self) -> None:                      # material/base.py:22-27
        self._sample_attributes()

    @WorldAssignmentWarning
    def set_world(self, _origin: torch.Tensor) -> None:
        super().set_world(_origin)

    @RelativeAssignmentWarning
    def setParent(self, parent) -> None:
        super().setParent(parent)

    @RelativeAssignmentWarning
    def setChild(self, child) -> None:
        super().setChild(child)

    @RotationAssignmentWarning
    def rotate_x(self, min_rot: float, max_rot: float) -> None:
        super().rotate_x(min_rot, max_rot)

    @RotationAssignmentWarning
    def rotate_y(self, min_rot: float, max_rot: float) -> None:
        super().rotate_y(min_rot, max_rot)

    @RotationAssignmentWarning
    def rotate_z(self, min_rot: float, max_rot: float) -> None:
        super().rotate_z(min_rot, max_rot)

    @RotationAssignmentWarning
    def rotate(self, min: torch.Tensor, max: torch.Tensor) -> None:
        super().rotate(min, max)

    @TranslationAssignmentWarning
    def translate_x(self, min_translation: float, max_translation: float) -> None:
        super().translate_x(min_translation, max_translation)

    @TranslationAssignmentWarning
    def translate_y(self, min_translation: float, max_translation: float) -> None:
        super().translate_y(min_translation, max_translation)

    @TranslationAssignmentWarning
    def translate_z(self, min_translation: float, max_translation: float) -> None:
        super().translate_z(min_translation, max_translation)

    @TranslationAssignmentWarning
    def translate(self, min: torch.Tensor, max: torch.Tensor) -> None:
        super().translate(min, max)

    @RotationAssignmentWarning
    def sample_rotation(self) -> torch.Tensor:
        return super().sample_rotation()

    @TranslationAssignmentWarning
    def sample_translation(self) -> torch.Tensor:
        return super().sample_translation()

    @RelativeAssignmentWarning
    def relative(self) -> bool:
        return super().relative()

    @WorldAssignmentWarning
    def world(self) -> torch.Tensor:
        return super().world()

    @WorldAssignmentWarning
    def nonRandomizedWorld(self) -> torch.Tensor:
        return super().nonRandomizedWorld()
