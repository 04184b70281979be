"""``fireflies/utils/intersections.py`` -- batched ray/plane and sphere/sphere tests on the device."""
import torch

from .. import _native as nat


def _rows(t: torch.Tensor, name: str, cols=None) -> torch.Tensor:
    t = nat.require_cuda(t.detach().float().contiguous(), torch.float32, name)
    if t.dim() != 2 or (cols is not None and t.shape[1] != cols):
        raise ValueError(f"{name}: expected a [N,{cols or 'D'}] tensor")
    return t


def rayPlane(laserOrigin, laserDirection, planeOrigin, planeNormal):
    """intersections.py:5-12: ray parameter of the hit, ``[N,1]``.  Operands broadcast to the rays' ``[N,3]``."""
    d = _rows(laserDirection, "laserDirection", 3)
    o, po, pn = (_rows(x.expand_as(d) if x.shape != d.shape else x, n, 3)
                 for x, n in ((laserOrigin, "laserOrigin"), (planeOrigin, "planeOrigin"), (planeNormal, "planeNormal")))
    t = torch.empty(d.shape[0], dtype=torch.float32, device=d.device)
    nat.check(nat.lib().ffb_ray_plane(o.data_ptr(), d.data_ptr(), po.data_ptr(), pn.data_ptr(), d.shape[0], t.data_ptr(), nat.stream()),
              "ffb_ray_plane")
    nat.count()
    return t[:, None]


def sphereSphere(a_coords, a_radius, b_coords, b_radius):
    """intersections.py:26-33: ``[N,1]`` bool, True where the spheres (circles, any dimension) touch or overlap."""
    a, b = _rows(a_coords, "a_coords"), _rows(b_coords, "b_coords")
    if a.shape != b.shape:
        raise ValueError("a_coords and b_coords must have the same shape")
    N = a.shape[0]
    ra = nat.require_cuda(torch.as_tensor(a_radius, device=a.device).detach().float().expand(N, 1).reshape(N).contiguous(), torch.float32, "a_radius")
    rb = nat.require_cuda(torch.as_tensor(b_radius, device=a.device).detach().float().expand(N, 1).reshape(N).contiguous(), torch.float32, "b_radius")
    hit = torch.empty(N, dtype=torch.uint8, device=a.device)
    nat.check(nat.lib().ffb_sphere_sphere(a.data_ptr(), ra.data_ptr(), b.data_ptr(), rb.data_ptr(), N, a.shape[1], hit.data_ptr(), nat.stream()),
              "ffb_sphere_sphere")
    nat.count()
    return hit.bool()[:, None]
