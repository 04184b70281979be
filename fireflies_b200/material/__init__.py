from .base import Material
