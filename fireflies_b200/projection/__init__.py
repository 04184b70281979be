from .camera import Camera
from .laser import Laser
