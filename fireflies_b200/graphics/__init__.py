from . import rasterization  # noqa: F401
