"""``fireflies/utils/io.py`` subset: config reader and the pytorch3d-style projection matrix
(utils/io.py:9-68) and the Blender NURBS .obj reader (utils/io.py:77-108; it returns this package's ``NurbsCurve``
instead of a geomdl ``NURBS.Curve``)."""
import math as _math
from pathlib import Path

import torch


def read_config_yaml(file_path: str) -> dict:
    import yaml
    return yaml.safe_load(Path(file_path).read_text())


def build_projection_matrix(fov: float, near_clip: float, far_clip: float,
                            device: torch.device = torch.device("cuda")) -> torch.Tensor:
    """utils/io.py:14-68: K mapping camera space to NDC [-1,1] (z_sign = -1).  Built on the host in fp32
    with the reference's operation order, then moved to ``device``."""
    K = torch.zeros((4, 4), dtype=torch.float32)
    t = torch.tan(torch.tensor((_math.pi / 180) * fov) / 2.0)
    max_y = t * near_clip
    min_y = -max_y
    max_x = max_y * 1.0
    min_x = -max_x
    K[0, 0] = 2.0 * near_clip / (max_x - min_x)
    K[1, 1] = 2.0 * near_clip / (max_y - min_y)
    K[0, 2] = (max_x + min_x) / (max_x - min_x)
    K[1, 2] = (max_y + min_y) / (max_y - min_y)
    K[3, 2] = -1.0
    K[2, 2] = -1.0 * far_clip / (far_clip - near_clip)
    K[2, 3] = -(far_clip * near_clip) / (far_clip - near_clip)
    return K.to(device)


def importBlenderNurbsObj(path, device: torch.device = torch.device("cuda")):
    """utils/io.py:77-108: control points from the ``v`` lines, the degree from ``deg``, the knot vector from ``parm u``."""
    from .nurbs import NurbsCurve
    control_points, deg, knotvector = [], None, None
    with open(path, "r") as obj_file:
        for line in obj_file:
            if line.startswith("v "):
                control_points.append([float(v) for v in line[2:].split()])
            elif line.startswith("deg "):
                deg = int(line[4:])
            elif line.startswith("parm u "):
                knotvector = [float(v) for v in line[7:].split()]
    return NurbsCurve(deg, control_points, knotvector, device=device)
