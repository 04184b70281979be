"""Warning decorators of ``fireflies/utils/warnings.py``.  The reference's Translation/World variants forget
to ``return new_func`` (so the decorated methods become ``None``, utils/warnings.py:39-66); here all four
return the wrapper -- the evident intent."""
import functools
import warnings


def _make(kind: str):
    def deco(func):
        @functools.wraps(func)
        def new_func(*args, **kwargs):
            warnings.simplefilter("always", Warning)
            warnings.warn(f"This object should generally not have a {kind} via {func.__name__}.", category=Warning, stacklevel=2)
            warnings.simplefilter("default", Warning)
            return func(*args, **kwargs)
        return new_func
    return deco


RotationAssignmentWarning = _make("transformation assignment")
RelativeAssignmentWarning = _make("parent/child assignment")
TranslationAssignmentWarning = _make("translation assignment")
WorldAssignmentWarning = _make("to-world matrix")
