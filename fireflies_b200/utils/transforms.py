"""``fireflies/utils/transforms.py`` is an EMPTY file in the reference although
``fireflies.utils.transforms.<fn>`` is called at 13 live sites; the evident intent
is ``fireflies.utils.math.<fn>``, which is what this module exports."""
from .math import transform_points, transform_directions, toMat4x4  # noqa: F401
