"""NURBS curve object behind ``fireflies.entity.Curve`` -- the stand-in for geomdl's ``NURBS.Curve`` that
``utils/io.py:105-108`` builds (``degree``, ``ctrlpts``, ``knotvector`` assigned in that order; optional ``weights``).

geomdl==5.3.1 is a third-party dependency absent from the reference tree and from this image, so parity with it is
unpinned; its conventions are followed as published: the knot vector is checked (``n + p + 1`` non-decreasing values)
and normalised to [0, 1] on assignment, control points without weights get weight 1, ``evaluate_single`` takes a
parameter in [0, 1] and returns Python floats.  Evaluation runs on the device in fp64 (``ffb_nurbs_curve_eval``).
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np
import torch

from .. import _native as nat

MAX_DEGREE = 7      # FFB_NURBS_MAX_DEGREE


class NurbsCurve:
    def __init__(self, degree: Optional[int] = None, ctrlpts: Optional[Sequence[Sequence[float]]] = None,
                 knotvector: Optional[Sequence[float]] = None, weights: Optional[Sequence[float]] = None,
                 device: torch.device = torch.device("cuda")):
        self._degree = None
        self._ctrlpts: List[List[float]] = []
        self._weights: Optional[List[float]] = None
        self._knots: List[float] = []
        self._device = device
        self._tables = None
        if degree is not None:
            self.degree = degree
        if ctrlpts is not None:
            self.ctrlpts = ctrlpts
        if weights is not None:
            self.weights = weights
        if knotvector is not None:
            self.knotvector = knotvector

    # ---- geomdl-style attributes -------------------------------------------------------------------------
    @property
    def degree(self) -> int:
        return self._degree

    @degree.setter
    def degree(self, value: int) -> None:
        value = int(value)
        if value < 1 or value > MAX_DEGREE:
            raise ValueError(f"degree must be in [1, {MAX_DEGREE}]")
        self._degree, self._tables = value, None

    @property
    def ctrlpts(self) -> List[List[float]]:
        return self._ctrlpts

    @ctrlpts.setter
    def ctrlpts(self, value) -> None:
        pts = [[float(c) for c in p] for p in value]
        if self._degree is None:
            raise ValueError("set the degree before the control points")
        if len(pts) < self._degree + 1:
            raise ValueError("a curve of degree p needs at least p + 1 control points")
        if any(len(p) != 3 for p in pts):
            raise ValueError("control points must be 3-dimensional")
        self._ctrlpts, self._tables = pts, None

    @property
    def weights(self) -> List[float]:
        return [1.0] * len(self._ctrlpts) if self._weights is None else self._weights

    @weights.setter
    def weights(self, value) -> None:
        w = [float(x) for x in value]
        if len(w) != len(self._ctrlpts) or any(x <= 0.0 for x in w):
            raise ValueError("one positive weight per control point")
        self._weights, self._tables = w, None

    @property
    def knotvector(self) -> List[float]:
        return self._knots

    @knotvector.setter
    def knotvector(self, value) -> None:
        kv = [float(k) for k in value]
        if self._degree is None or not self._ctrlpts:
            raise ValueError("set the degree and the control points before the knot vector")
        if len(kv) != len(self._ctrlpts) + self._degree + 1:
            raise ValueError("the knot vector needs len(ctrlpts) + degree + 1 values")
        if any(b < a for a, b in zip(kv, kv[1:])) or kv[-1] <= kv[0]:
            raise ValueError("the knot vector must be non-decreasing and span a positive interval")
        first, last = kv[0], kv[-1]
        self._knots = [float("{:.18f}".format((k - first) / (last - first))) for k in kv]
        self._tables = None

    # ---- device tables ------------------------------------------------------------------------------------
    def tables(self):
        """``(ctrlw f64 [n,4], knots f64 [n+p+1])`` on the device, rebuilt when an attribute is reassigned."""
        if self._tables is None:
            if self._degree is None or not self._ctrlpts or not self._knots:
                raise ValueError("degree, ctrlpts and knotvector must be set before evaluating")
            w = np.asarray(self.weights, dtype=np.float64)
            cw = np.concatenate([np.asarray(self._ctrlpts, dtype=np.float64) * w[:, None], w[:, None]], axis=1)
            self._tables = (torch.from_numpy(cw).to(self._device), torch.tensor(self._knots, dtype=torch.float64, device=self._device))
        return self._tables

    def _params(self, t, margin: float = 0.0) -> torch.Tensor:
        """Host values are range-checked like geomdl does (it raises outside the domain); a device tensor is taken as is
        (checking it would cost a host sync: parameters outside [0, 1] then extrapolate the end spans)."""
        if torch.is_tensor(t):
            p = t.detach().to(self._device, torch.float64).reshape(-1).contiguous()
        else:
            vals = [float(v) for v in np.atleast_1d(np.asarray(t, dtype=np.float64))]
            if any(not (0.0 <= v and v + margin <= 1.0) for v in vals):
                raise ValueError("curve parameter outside [0, 1]")
            p = torch.tensor(vals, dtype=torch.float64, device=self._device)
        return nat.require_cuda(p, torch.float64, "t")

    def evaluate(self, t) -> torch.Tensor:
        """Curve points for a batch of parameters: f64 ``[B,3]`` on the device (one launch)."""
        cw, kn = self.tables()
        p = self._params(t)
        out = torch.empty((p.shape[0], 3), dtype=torch.float64, device=p.device)
        nat.check(nat.lib().ffb_nurbs_curve_eval(cw.data_ptr(), kn.data_ptr(), cw.shape[0], self._degree, p.data_ptr(), p.shape[0],
                                                 out.data_ptr(), nat.stream()), "ffb_nurbs_curve_eval")
        nat.count()
        return out

    def evaluate_single(self, param: float) -> List[float]:
        """geomdl's ``evaluate_single``: one parameter in, a list of three Python floats out (host sync)."""
        return self.evaluate([param])[0].tolist()

    def evaluate_list(self, params) -> List[List[float]]:
        return self.evaluate(list(params)).tolist()

    def poses(self, t, world: torch.Tensor, dt: float = 0.001, parts: bool = False):
        """``Curve.randomize``'s matrix for a batch of path parameters (entity/curve.py:48-96), one launch:
        ``T(C(t)) @ toMat4x4(R([0,1,0] -> C(t+dt) - C(t), x/z negated)) @ world`` -> f32 ``[B,4,4]``.
        ``parts=True`` also returns the rotation and translation matrices."""
        cw, kn = self.tables()
        p = self._params(t, margin=dt)
        W = nat.require_cuda(world.detach().to(p.device).float().contiguous(), torch.float32, "world")
        B = p.shape[0]
        out = torch.empty((B, 4, 4), dtype=torch.float32, device=p.device)
        rot = torch.empty_like(out) if parts else None
        tr = torch.empty_like(out) if parts else None
        nat.check(nat.lib().ffb_curve_pose(cw.data_ptr(), kn.data_ptr(), cw.shape[0], self._degree, p.data_ptr(), B, float(dt),
                                           W.data_ptr(), out.data_ptr(), nat.ptr(rot), nat.ptr(tr), nat.stream()), "ffb_curve_pose")
        nat.count()
        return (out, rot, tr) if parts else out


def as_nurbs(curve, device) -> NurbsCurve:
    """Accept a ``NurbsCurve`` or any geomdl-like object exposing ``degree``, ``ctrlpts``, ``knotvector`` (and ``weights``)."""
    if isinstance(curve, NurbsCurve):
        return curve
    weights = getattr(curve, "weights", None)
    return NurbsCurve(curve.degree, curve.ctrlpts, curve.knotvector, weights if weights else None, device=device)
