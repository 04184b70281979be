"""fireflies_b200 -- B200 (sm_100a) implementation of the Fireflies hot path.

Drop-in for the part of Henningson/Fireflies the project owns: ``import fireflies_b200 as fireflies``.
All compute runs in ``libffb200.so`` (hand-written CUDA, C ABI in ``include/ffb200.h``); there is no CPU
fallback -- importing works anywhere, calling a kernel without the library or a GPU raises.
"""
from . import _native  # noqa: F401
from . import graphics  # noqa: F401

__version__ = "0.1.0"
