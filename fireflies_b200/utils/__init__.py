from . import math  # noqa: F401
from . import warnings  # noqa: F401
from . import io  # noqa: F401
from . import transforms  # noqa: F401
