"""A stand-in for the ``mi.traverse(scene)`` parameter map: ``keys()``,
``__getitem__/__setitem__``, ``update()``, ``*.to_world`` values exposing ``.matrix.torch()`` -> [1,4,4], and the
value types ``Scene`` touches.  Works for both the reference (through oracle/ref_loader.py) and fireflies_b200."""
import torch


class _Matrix:
    def __init__(self, m):
        self._m = torch.as_tensor(m, dtype=torch.float32).reshape(1, 4, 4)

    def torch(self):
        return self._m.clone()


class Transform4f:
    def __init__(self, m):
        self.matrix = _Matrix(m)


class ScalarTransform3f:
    def __init__(self, m=None):
        self.m = m


class Float(float):
    is_scalar = True


class Float32(list):
    def __init__(self, data):
        super().__init__(torch.as_tensor(data).flatten().tolist())

    def torch(self):
        return torch.tensor(list(self), dtype=torch.float32)


class Color3f(list):
    def torch(self):
        return torch.tensor(list(self), dtype=torch.float32).reshape(1, 3)


class TensorXf:
    def __init__(self, t):
        self.t = torch.as_tensor(t)

    def torch(self):
        return self.t


class Types:
    Float32, Transform4f, TensorXf, Float, ScalarTransform3f = Float32, Transform4f, TensorXf, Float, ScalarTransform3f


class FakeParams(dict):
    types = Types
    n_updates = 0

    def update(self, *a, **k):      # mi.SceneParameters.update()
        if a or k:
            return super().update(*a, **k)
        self.n_updates += 1


def install_into(mi_module):
    """Give a stub ``mitsuba`` module (oracle/ref_loader.py) the fake value types."""
    for n in ("Float32", "Transform4f", "TensorXf", "Float", "ScalarTransform3f"):
        setattr(mi_module, n, getattr(Types, n))


def demo_params(seed: int = 0, n_a: int = 301, n_b: int = 77) -> FakeParams:
    g = torch.Generator().manual_seed(seed)
    p = FakeParams()
    p["mesh-A.vertex_positions"] = Float32(torch.rand(n_a * 3, generator=g) * 2 - 1)
    p["mesh-B.vertex_positions"] = Float32(torch.rand(n_b * 3, generator=g) + 2)
    W = torch.eye(4)
    W[:3, 3] = torch.tensor([0.1, 0.2, 3.0])
    p["PerspectiveCamera.to_world"] = Transform4f(W)
    p["PerspectiveCamera.x_fov"] = Float(45.0)
    p["Projector.to_world"] = Transform4f(torch.eye(4))
    p["emit-Spot.to_world"] = Transform4f(W.clone())
    p["emit-Spot.intensity.value"] = Color3f([1.0, 2.0, 3.0])
    p["emit-Spot.cutoff_angle"] = Float(20.0)
    p["mat-Mucosa.brdf_0.roughness.value"] = Float(0.5)
    p["mat-Mucosa.brdf_0.base_color.value"] = Color3f([0.2, 0.3, 0.4])
    return p
