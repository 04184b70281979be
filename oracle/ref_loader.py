"""TEST INFRASTRUCTURE ONLY -- imports the *real* reference (pure Python/PyTorch)
from ``/root/reference`` with empty stub modules for its absent third-party
imports (mitsuba, drjit, kornia, geomdl, pywavefront), following SURVEY.md
Appendix C.  Only usable in the build container: ``/root/reference`` does not
exist on the GPU box, so nothing in the ``-m gpu`` tests, ``smoke()`` or
``bench.py`` may call this.  It is used by ``oracle/make_golden.py`` (to write
``tests/golden/*.npz``) and by the ``not gpu`` tests that cross-check the
restatement in ``oracle/ff_oracle.py`` against the reference when it is present.
"""
import os
import sys
import types

REF_ROOT = os.environ.get("FIREFLIES_REF", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "fireflies", "__init__.py"))


def load():
    """Return the reference ``fireflies`` package (imported once, cached)."""
    if "fireflies" in sys.modules and getattr(sys.modules["fireflies"], "_ffb_ref", False):
        return sys.modules["fireflies"]
    if not available():
        raise RuntimeError(f"reference not found under {REF_ROOT}")
    for name in ("mitsuba", "drjit", "kornia", "geomdl", "pywavefront"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["geomdl"].NURBS = types.SimpleNamespace(Curve=object)
    sys.modules["drjit"].wrap_ad = lambda **kw: (lambda f: f)
    kf = types.ModuleType("kornia.filters")
    sys.modules["kornia"].filters = kf
    sys.modules["kornia.filters"] = kf
    # mitsuba types touched by scene.py (fakes; see tests/fake_mitsuba.py for the params object)
    mi = sys.modules["mitsuba"]
    for tname in ("Float", "Float32", "Transform4f", "ScalarTransform3f", "TensorXf"):
        if not hasattr(mi, tname):
            setattr(mi, tname, type(tname, (), {}))
    sys.path.insert(0, REF_ROOT)
    try:
        import fireflies  # noqa: F401
        import fireflies.graphics.rasterization  # noqa: F401
        import fireflies.projection  # noqa: F401
        import fireflies.postprocessing  # noqa: F401
        import fireflies.utils.math  # noqa: F401
    finally:
        sys.path.remove(REF_ROOT)
    ff = sys.modules["fireflies"]
    # intent shims for reference bugs (fireflies/utils/transforms.py is an empty file)
    tr = sys.modules["fireflies.utils.transforms"]
    m = sys.modules["fireflies.utils.math"]
    for fn in ("transform_points", "transform_directions", "toMat4x4"):
        setattr(tr, fn, getattr(m, fn))
    ff._ffb_ref = True
    return ff
