from .base import Transformable
from .mesh import Mesh
from .curve import Curve
