from .base import Light
