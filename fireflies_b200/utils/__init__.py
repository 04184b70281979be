from . import math  # noqa: F401
from . import warnings  # noqa: F401
from . import io  # noqa: F401
from . import transforms  # noqa: F401
from . import intersections  # noqa: F401
from . import torch_grads  # noqa: F401
from . import nurbs  # noqa: F401
