"""GPU parity of the NURBS-curve camera path (fireflies/entity/curve.py), the blue-noise ray generator
(projection/laser.py:94-145) and the batched intersections (utils/intersections.py), all through the C ABI.

Bars: the fp64 curve points are bit-exact against the oracle's restatement of the evaluator (same IEEE operations in
the same order; geomdl itself is absent: parity with it is unpinned) and within 1e-14 of closed forms; pose matrices
within 1e-5 relative (+1e-6 absolute) of the reference's Curve methods (tests/golden/curve.npz); path parameters, hit
masks and sample counts exact."""
import math
import random

import numpy as np
import pytest
import torch

from oracle import ff_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ff():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import fireflies_b200 as ff
    return ff


def close(a, b, rtol=1e-5, atol=1e-6):
    a = a.detach().cpu().double().numpy() if torch.is_tensor(a) else np.asarray(a, np.float64)
    b = b.detach().cpu().double().numpy() if torch.is_tensor(b) else np.asarray(b, np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    err = np.abs(a - b) - (atol + rtol * np.abs(b))
    assert err.max() <= 0, f"max violation {err.max():.3e}; max abs diff {np.abs(a - b).max():.3e}"


def make_curve(ff, g, tag):
    return ff.utils.nurbs.NurbsCurve(3, g["ctrl"].tolist(), g["knots"].tolist(), g["weights"].tolist() if tag == "rational" else None)


@pytest.mark.parametrize("tag", ["bspline", "rational"])
def test_curve_points_bit_exact_fp64(ff, golden, tag):
    g = golden("curve")
    c = make_curve(ff, g, tag)
    ts = np.linspace(0.0, 1.0, 101)
    pts = c.evaluate(ts).cpu().numpy()
    assert pts.dtype == np.float64 and np.array_equal(pts, g[tag + "_points64"])
    assert c.evaluate_single(0.37) == O.nurbs_curve_point(g["ctrl"], c.knotvector, 3, 0.37, g["weights"] if tag == "rational" else None)
    assert c.evaluate(torch.tensor(ts, device="cuda")).equal(torch.from_numpy(pts).cuda())      # device parameters, no host check
    assert c.evaluate([]).shape == (0, 3)


def test_curve_known_answers_and_degrees(ff):
    N = ff.utils.nurbs.NurbsCurve
    ctrl = [[0, 0, 0], [1, 2, 0], [3, 2, 1], [4, 0, 2]]
    bez = N(3, ctrl, [0, 0, 0, 0, 2, 2, 2, 2])
    ts = np.linspace(0, 1, 33)
    b = np.stack([(1 - ts) ** 3, 3 * ts * (1 - ts) ** 2, 3 * ts * ts * (1 - ts), ts ** 3], 1)
    close(bez.evaluate(ts), b @ np.array(ctrl, np.float64), rtol=1e-14, atol=1e-15)
    arc = N(2, [[1, 0, 0], [1, 1, 0], [0, 1, 0]], [0, 0, 0, 1, 1, 1], [1, math.sqrt(0.5), 1])
    p = arc.evaluate(ts).cpu().numpy()
    assert np.abs(p[:, 0] ** 2 + p[:, 1] ** 2 - 1).max() < 1e-14 and np.all(p[:, 2] == 0)
    rng = np.random.default_rng(7)
    for deg in (1, 2, 5, 7):                      # every supported degree, random non-uniform knots, against the oracle
        n = deg + 6
        cp = rng.normal(size=(n, 3)).tolist()
        w = (rng.random(n) + 0.5).tolist()
        kn = [0.0] * (deg + 1) + sorted(rng.random(n - deg - 1).tolist()) + [1.0] * (deg + 1)
        c = N(deg, cp, kn, w)
        got = c.evaluate(ts).cpu().numpy()
        want = np.array([O.nurbs_curve_point(cp, c.knotvector, deg, float(t), w) for t in ts])
        assert np.array_equal(got, want), deg
    with pytest.raises(ValueError):
        N(8, [[0, 0, 0]] * 10, list(range(19)))


@pytest.mark.parametrize("tag", ["bspline", "rational"])
def test_curve_entity_walk_matches_reference(ff, golden, tag):
    g = golden("curve")
    W = torch.from_numpy(g[tag + "_W"]).cuda()
    deltas, worlds = g[tag + "_deltas"], g[tag + "_worlds"]
    c = ff.entity.Curve("path", make_curve(ff, g, tag))
    c.set_world(W)
    got_d, got_w = [], []
    c.train()
    random.seed(5)
    for _ in range(3):
        c.randomize(); got_d.append(c.curve_delta); got_w.append(c.world())
    c.eval()
    for _ in range(40):
        c.randomize(); got_d.append(c.curve_delta); got_w.append(c.world())
    c.curve_delta = 0.9485
    for _ in range(4):
        c.randomize(); got_d.append(c.curve_delta); got_w.append(c.world())
    assert np.array_equal(np.array(got_d), deltas)                   # path parameters: exact
    close(torch.stack(got_w), worlds)
    for t in (0.2, 0.5, 0.77):
        c.curve_delta = t
        close(c.sample_rotation(), g[f"{tag}_rot_{t}"])
        close(c.sample_translation(), g[f"{tag}_trans_{t}"])
    # one launch for a whole batch of steps == the same steps one by one
    a = ff.entity.Curve("a", make_curve(ff, g, tag)); a.set_world(W); a.eval()
    b = ff.entity.Curve("b", make_curve(ff, g, tag)); b.set_world(W); b.eval()
    batch = a.randomize_batch(1200)                                  # crosses the wrap at 1 - epsilon
    seq = []
    for _ in range(1200):
        b.randomize(); seq.append(b.world())
    assert torch.equal(batch, torch.stack(seq)) and a.curve_delta == b.curve_delta and torch.equal(a.world(), b.world())
    # a parented curve camera: world() chains like any Transformable
    parent = ff.entity.Transformable("rig"); parent.set_world(torch.eye(4, device="cuda") * 2.0)
    b.setParent(parent)
    close(b.world(), (torch.eye(4) * 2.0) @ seq[-1].cpu())


def test_curve_pose_oracle_sweep(ff, golden):
    g = golden("curve")
    c = make_curve(ff, g, "rational")
    kn = c.knotvector
    W = torch.eye(4)
    W[0:3, 3] = torch.tensor([1.0, -2.0, 0.5])
    ts = np.linspace(0.0, 0.999, 257)
    got, rot, tr = c.poses(ts, W.cuda(), parts=True)
    want = torch.stack([O.curve_pose(g["ctrl"], kn, 3, float(t), W, g["weights"]) for t in ts])
    close(got, want)
    close(tr[:, 0:3, 3], c.evaluate(ts).float())
    rr = rot[:, 0:3, 0:3].double()
    close(rr @ rr.transpose(1, 2), torch.eye(3).expand(257, 3, 3), rtol=0, atol=1e-3)      # Rodrigues: orthonormal (fp32, tiny tangents)
    with pytest.raises(ValueError):
        c.poses([0.9995], W.cuda())                                   # t + dt leaves the domain: geomdl raises


def test_blue_noise_rays_and_misc(ff, golden):
    g = golden("poisson_misc")
    K = torch.from_numpy(g["blue_K"]).cuda()
    np.random.seed(19)
    rays = ff.projection.Laser.generate_blue_noise_rays(64, 48, 60, K)
    assert rays.shape == g["blue_rays"].shape
    close(rays, g["blue_rays"])
    I = ff.utils.intersections
    t = I.rayPlane(*[torch.from_numpy(g[k]).cuda() for k in ("rp_o", "rp_d", "rp_po", "rp_pn")])
    assert t.shape == (33, 1)
    close(t, g["rp_t"], rtol=1e-5, atol=1e-6)
    o = torch.zeros(4, 3, device="cuda")
    d = torch.tensor([[0.0, 0.0, 1.0], [1.0, 0.0, 0.0], [1.0, 0.0, 1e-7], [0.0, 1.0, -2.0]], device="cuda")
    po, pn = torch.tensor([[0.0, 0.0, 5.0]], device="cuda").expand(4, 3), torch.tensor([[0.0, 0.0, 1.0]], device="cuda").expand(4, 3)
    t = I.rayPlane(o, d, po, pn).cpu()
    want = O.ray_plane(o.cpu(), d.cpu(), po.cpu(), pn.cpu())
    assert np.array_equal(t.numpy(), want.numpy(), equal_nan=True) and math.isnan(t[1, 0]) and t[2, 0] == 5.0 and t[0, 0] == 5.0
    hit = I.sphereSphere(*[torch.from_numpy(g[k]).cuda() for k in ("ss_a", "ss_ra", "ss_b", "ss_rb")])
    assert hit.dtype == torch.bool and np.array_equal(hit.cpu().numpy(), g["ss_hit"])
    a3 = torch.tensor([[0.0, 0.0, 0.0], [0.0, 0.0, 0.0]], device="cuda")
    b3 = torch.tensor([[3.0, 4.0, 0.0], [3.0, 4.0, 0.1]], device="cuda")
    assert I.sphereSphere(a3, torch.tensor([[2.0], [2.0]], device="cuda"), b3, torch.tensor([[3.0], [3.0]], device="cuda")).flatten().tolist() == [True, False]
    M = ff.utils.math
    v1, v2 = torch.from_numpy(g["rot_v1"]).cuda(), torch.from_numpy(g["rot_v2"]).cuda()
    close(M.rotation_matrix_from_vectors(v1, v2), g["rot"])
    close(M.rotation_matrix_from_vectors_with_fixed_up(v1, v2), g["rot_up"])


def test_rand_perlin_functions_consume_torch_rand_like_the_reference(ff, golden):
    import fireflies_b200.sampling.noise_texture_lerp as NT
    g = golden("perlin")
    shape, res, octaves = g["a_cfg"][:2].tolist(), g["a_cfg"][2:4].tolist(), int(g["a_cfg"][4])
    torch.manual_seed(21)
    noise = NT.rand_perlin_2d_octaves(shape, res, octaves, float(g["a_pers"]))
    close(noise, g["a_noise"], rtol=1e-5, atol=1e-5 * float(np.abs(g["a_noise"]).max()))
    torch.manual_seed(21)
    one = NT.rand_perlin_2d(shape, res)
    torch.manual_seed(21)
    close(one, NT.rand_perlin_2d_octaves(shape, res, 1, 0.5), rtol=0, atol=0)
