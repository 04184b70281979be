"""GPU parity: sampling, entity compose, vertex transforms, laser glue and the Scene facade, through the C ABI,
against the CPU oracle and the reference-generated fixtures.  fp32 outputs: 1e-5 relative (+ an absolute floor
tied to the magnitude of the inputs, since transformed coordinates can cancel to ~0); eval / integer
sequences: bit-exact."""
import random

import numpy as np
import pytest
import torch

from oracle import ff_oracle as O
import fake_mitsuba as fm

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ff():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import fireflies_b200
    return fireflies_b200


def T(a):
    return torch.from_numpy(np.asarray(a))


def close(a, b, rtol=1e-5, atol=1e-6):
    a = a.detach().cpu().double().numpy() if torch.is_tensor(a) else np.asarray(a, np.float64)
    b = b.detach().cpu().double().numpy() if torch.is_tensor(b) else np.asarray(b, np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    err = np.abs(a - b) - (atol + rtol * np.abs(b))
    assert err.max() <= 0, f"max violation {err.max():.3e}; max abs diff {np.abs(a - b).max():.3e}"


def test_eval_samplers_bit_exact(ff, golden):
    g = golden("samplers")
    S = ff.sampling
    us = S.UniformSampler(torch.zeros(3).cuda(), torch.zeros(3).cuda())
    us.get_min()[2], us.get_max()[2] = -np.pi, np.pi
    us.eval()
    assert np.array_equal(np.stack([us.sample().cpu().numpy() for _ in range(12)]), g["eval_vec3"])
    sc = S.UniformSampler(0.0, 0.05)
    sc.eval()
    assert np.array_equal(np.stack([sc.sample().cpu().numpy() for _ in range(14)]), g["eval_scalar"])
    v3 = S.UniformSampler(torch.tensor([0.0, 1.0, -1.0]).cuda(), torch.tensor([0.035, 1.5, 0.0]).cuda())
    v3.eval()
    assert np.array_equal(np.stack([v3.sample().cpu().numpy() for _ in range(12)]), g["eval_vec3_ranged"])
    s2v = S.UniformScalarToVec3Sampler(0.1, 10.0)
    s2v.eval()
    assert np.array_equal(np.stack([s2v.sample().cpu().numpy() for _ in range(4)]), g["eval_s2v"])
    an = S.AnimationSampler(0, 5, 0, 5)
    an.eval()
    assert [an.sample() for _ in range(9)] == g["anim_eval"].tolist()
    an.train()
    random.seed(1)
    assert [an.sample() for _ in range(5)] == g["anim_train_seed1"].tolist()


def test_train_uniform_injected_bit_exact(ff, golden):
    """u*(b-a)+a with the reference's torch.rand variates injected -> identical bits."""
    from fireflies_b200.sampling.base import _lerp_native
    g = golden("samplers")
    mn, mx = torch.tensor([-1.0, 0.0, 2.0]).cuda(), torch.tensor([1.0, 0.5, 2.0]).cuda()
    got = np.stack([_lerp_native(mn, mx, T(u).cuda()).cpu().numpy() for u in g["train_uniform_u"]])
    assert np.array_equal(got, g["train_uniform"])
    # and the train-mode sampler draws from torch's global generator, one rand(shape) per sample like the reference
    s = ff.sampling.UniformSampler(mn, mx)
    torch.manual_seed(3)
    a = s.sample()
    torch.manual_seed(3)
    u = torch.rand(3, device="cuda")
    assert torch.equal(a, _lerp_native(mn, mx, u))


def test_mesh_kat2_and_seeded_randomize(ff, golden):
    g = golden("transforms")
    Mesh, US = ff.entity.Mesh, ff.sampling.UniformSampler
    m = Mesh("m", T(g["kat2_verts"]).cuda())
    m.set_centroid(torch.tensor([[0.5, -0.5, 2.0]]))
    c = lambda v: torch.tensor(v).cuda()  # noqa: E731
    m.set_rotation_sampler(US(c([0.1, 0.2, 0.3]), c([0.1, 0.2, 0.3])))
    m.set_translation_sampler(US(c([1.0, 2.0, 3.0]), c([1.0, 2.0, 3.0])))
    m.set_scale_sampler(US(c([2.0, 1.0, 0.5]), c([2.0, 1.0, 0.5])))
    m.set_randomizable(True)
    m.train()
    m.randomize()
    close(m.world(), g["kat2_world"], rtol=1e-5, atol=1e-6)
    close(m.get_randomized_vertices(), g["kat2_out"], rtol=1e-5, atol=1e-5)
    # ranged T/R/S, non-identity world: inject the reference's variates through the batched kernels
    verts = T(g["rand_verts"]).cuda()
    m = Mesh("r", verts)
    m.set_world(T(g["rand_world0"]).cuda())
    m.set_centroid(T(g["rand_centroid"]).reshape(1, 3))
    m.rotate(c([-0.5, -1.0, -3.0]), c([0.5, 1.0, 3.0]))
    m.translate(c([-1.0, -2.0, -3.0]), c([1.0, 2.0, 3.0]))
    m.scale(c([0.5, 0.8, 1.0]), c([2.0, 1.2, 3.0]))
    params = fm.FakeParams()
    sc = ff.Scene(params)
    sc._meshes.append(m)
    sb = sc.batch(seed=1)
    res = sb.randomize(8, variates=T(g["rand_u"]).cuda())
    close(res.world[:, 0], g["rand_worlds"], rtol=1e-5, atol=1e-5)
    close(res.mesh_vertices("r"), g["rand_outs"], rtol=1e-5, atol=2e-5)
    # per-object API draws translation, rotation, scale from torch's generator in that order
    torch.manual_seed(77)
    m.train()
    m.randomize()
    torch.manual_seed(77)
    u = torch.stack([torch.rand(3, device="cuda") for _ in range(3)]).reshape(1, 3, 3)
    close(m.world(), sb.randomize(1, variates=u).world[0, 0], rtol=1e-6, atol=1e-6)


def test_parent_child_eval_chain(ff, golden):
    g = golden("transforms")
    Mesh = ff.entity.Mesh
    a, b = Mesh("a", T(g["chain_va"]).cuda()), Mesh("b", T(g["chain_vb"]).cuda())
    a.set_centroid(torch.tensor([[1.0, 0.0, 0.0]]))
    b.set_centroid(torch.tensor([[0.0, 2.0, 0.0]]))
    b.setParent(a)
    b.set_randomizable(True)
    a.rotate_z(-np.pi, np.pi)
    a.eval(); b.eval()
    for i in range(g["chain_rot"].shape[0]):
        a.randomize(); b.randomize()
        assert np.array_equal(a._sampled_rotation.cpu().numpy(), g["chain_rot"][i])       # eval sequence: bit-exact
        close(a.world(), g["chain_worlds"][i, 0], rtol=1e-5, atol=1e-6)
        close(b.world(), g["chain_worlds"][i, 1], rtol=1e-5, atol=1e-5)
        close(b.get_randomized_vertices(), g["chain_child_verts"][i], rtol=1e-5, atol=1e-5)
    # same sequence through the batched path (fresh objects): 5 eval steps in ONE launch
    a2, b2 = Mesh("a", T(g["chain_va"]).cuda()), Mesh("b", T(g["chain_vb"]).cuda())
    a2.set_centroid(torch.tensor([[1.0, 0.0, 0.0]])); b2.set_centroid(torch.tensor([[0.0, 2.0, 0.0]]))
    b2.setParent(a2); b2.set_randomizable(True); a2.rotate_z(-np.pi, np.pi)
    sc = ff.Scene(fm.FakeParams())
    sc._meshes += [b2, a2]            # child listed first: the batch must still order parents first
    sc.eval()
    res = sc.batch().randomize(5)
    close(res.entity_world("a"), g["chain_worlds"][:, 0], rtol=1e-5, atol=1e-6)
    close(res.entity_world("b"), g["chain_worlds"][:, 1], rtol=1e-5, atol=1e-5)
    close(res.mesh_vertices("b"), g["chain_child_verts"], rtol=1e-5, atol=1e-5)


def test_batch_orders_parents_across_lists_and_trees(ff):
    """Parents in another entity list (a light and a mesh parented to the camera) and a parent with two children: the batched world
    matrices equal Transformable.world() of the same objects (reference: world = parent.world() @ local, entity/base.py:239-244)."""
    c = lambda v: torch.tensor(v, dtype=torch.float32).cuda()  # noqa: E731
    g = torch.Generator().manual_seed(11)
    cam = ff.entity.Transformable("PerspectiveCamera")
    cam.set_world(c([[0.0, -1.0, 0.0, 1.0], [1.0, 0.0, 0.0, 2.0], [0.0, 0.0, 1.0, 3.0], [0.0, 0.0, 0.0, 1.0]]))
    cam.rotate_x(0.2, 0.2); cam.translate_y(0.5, 0.5)                 # degenerate ranges: the "random" pose is known
    light = ff.entity.Transformable("light-on-camera")
    light.set_world(c([[1.0, 0.0, 0.0, 0.0], [0.0, 1.0, 0.0, 0.25], [0.0, 0.0, 1.0, -0.5], [0.0, 0.0, 0.0, 1.0]]))
    light.setParent(cam); light.rotate_y(-0.4, -0.4)
    p = ff.entity.Mesh("tree-parent", torch.rand(8, 3, generator=g).cuda())
    c1 = ff.entity.Mesh("tree-child-1", torch.rand(8, 3, generator=g).cuda())
    c2 = ff.entity.Mesh("tree-child-2", torch.rand(8, 3, generator=g).cuda())
    g1 = ff.entity.Mesh("tree-grandchild", torch.rand(8, 3, generator=g).cuda())
    rig = ff.entity.Mesh("mesh-on-camera", torch.rand(8, 3, generator=g).cuda())
    p.rotate_z(0.7, 0.7); p.translate_x(-1.0, -1.0)
    for ch, par, ang in ((c1, p, 0.3), (c2, p, -0.6), (g1, c1, 1.1), (rig, cam, 0.15)):
        ch._parent = par                                              # setParent keeps ONE child slot; a tree needs the bare link
        ch.set_randomizable(True); ch.rotate_y(ang, ang); ch.scale_x(1.5, 1.5)
    sc = ff.Scene(fm.FakeParams())
    sc._meshes += [g1, c1, c2, p, rig]                                # worst case: every child listed before its parent
    sc._lights.append(light)
    sc._camera = cam
    sc.train()
    sb = sc.batch(seed=5)
    order = {e.name(): i for i, e in enumerate(sb.entities)}
    for e in sb.entities:
        if e.parent() is not None:
            assert order[e.parent().name()] < order[e.name()]
    res = sb.randomize(2)
    for e in (cam, p, c1, c2, g1, rig, light):                        # parents first: world() reads the parent's randomised matrix
        e.randomize()
    for e in (cam, light, p, c1, c2, g1, rig):
        close(res.entity_world(e.name())[1], e.world(), rtol=1e-5, atol=1e-5)
    close(res.mesh_vertices("tree-grandchild")[0], g1.get_randomized_vertices(), rtol=1e-5, atol=1e-5)
    # a parent outside the scene, or re-parenting after the batch was built, is an error instead of a silently dropped parent
    stray = ff.entity.Transformable("stray")
    c2._parent = stray
    with pytest.raises(ValueError):
        sc.batch()
    c2._parent = p
    c1._parent = c2                                                   # c2's row follows c1's in the existing table
    with pytest.raises(ValueError):
        sb.refresh()


def test_transformable_attributes(ff, golden):
    g = golden("transforms")
    c = lambda v: torch.tensor(v).cuda()  # noqa: E731
    t = ff.entity.Transformable("light")
    t.set_world(T(g["tr_world0"]).cuda())
    t.rotate_x(-0.3, 0.3)
    t.translate_y(-1.0, 1.0)
    t.add_float_key("power", 1.0, 3.0)
    t.add_vec3_key("color", c([0.0, 0.1, 0.2]), c([1.0, 0.9, 0.8]))
    t.add_vec3_sampler("intensity", ff.sampling.UniformScalarToVec3Sampler(0.1, 10.0))
    sc = ff.Scene(fm.FakeParams())
    sc._lights.append(t)
    sb = sc.batch()
    u = T(g["tr_u"])                     # per call: translation(3) rotation(3) power(1) color(3) intensity(1)
    v = torch.zeros(4, sb.S, 3)
    v[:, 0], v[:, 1] = u[:, 0:3], u[:, 3:6]
    v[:, sb._attr_rows[("light", "power")][0], 0] = u[:, 6]
    v[:, sb._attr_rows[("light", "color")][0]] = u[:, 7:10]
    v[:, sb._attr_rows[("light", "intensity")][0], 0] = u[:, 10]
    res = sb.randomize(4, variates=v.cuda())
    close(res.entity_world("light"), g["tr_worlds"], rtol=1e-5, atol=1e-5)
    assert np.array_equal(res.attribute("light", "power").cpu().numpy(), g["tr_power"])
    assert np.array_equal(res.attribute("light", "color").cpu().numpy(), g["tr_color"])
    assert np.array_equal(res.attribute("light", "intensity").cpu().numpy(), g["tr_intensity"])


def test_transform_points_and_autograd(ff, golden):
    g = golden("transforms")
    M = ff.utils.math
    pts, K = T(g["tp_pts"]).cuda(), T(g["tp_K"]).cuda()
    close(M.transform_points(pts, K), g["tp_out"], rtol=1e-5, atol=1e-5)
    close(M.transform_directions(pts, T(g["rand_world0"]).cuda()), g["td_out"], rtol=1e-5, atol=1e-6)
    p = pts.clone().requires_grad_(True)
    w = torch.randn(pts.shape, generator=torch.Generator().manual_seed(1)).cuda()
    (M.transform_points(p, K) * w).sum().backward()
    pr = pts.cpu().clone().requires_grad_(True)
    (O.transform_points(pr, K.cpu()) * w.cpu()).sum().backward()
    close(p.grad, pr.grad, rtol=1e-4, atol=1e-4 * pr.grad.abs().max().item())
    # sizes that exercise the vector / scalar / tail paths
    for V in (1, 3, 4, 5, 1023, 1024, 1025, 4099):
        x = torch.rand(V, 3, generator=torch.Generator().manual_seed(V)) * 2 - 1
        close(M.transform_points(x.cuda(), K), O.transform_points(x, K.cpu()), rtol=1e-5, atol=1e-5)


def test_laser(ff, golden):
    g = golden("laser")
    Laser = ff.projection.Laser
    rays = Laser.generate_uniform_rays(0.0275, 18, 18)
    assert np.array_equal(rays.cpu().numpy(), g["rays"])
    K = ff.utils.io.build_projection_matrix(60, 0.01, 1000.0)
    close(K, g["K"], rtol=1e-6, atol=0)
    tr = ff.entity.Transformable("projector")
    laser = Laser(tr, rays, K, 60.0, 0.01, 1000.0)
    ndc = laser.projectRaysToNDC()
    close(ndc, g["ndc"], rtol=1e-5, atol=1e-6)
    close(laser.projectNDCPointsToWorld(ndc), g["back"], rtol=1e-4, atol=1e-5)
    laser2 = Laser(tr, T(g["wide"]).cuda(), T(g["K01"]).cuda(), 60.0, 0.01, 1000.0)
    laser2.clamp_to_fov()
    close(laser2._rays, g["wide_clamped"], rtol=1e-5, atol=1e-6)
    tex = laser.generateTexture(10.0, torch.tensor([64, 48]))
    assert tex.shape == (324, 48, 64) and tex.is_cuda
    close(tex.sum(0), g["gen_tex_sum"], rtol=1e-5, atol=1e-6)
    close(laser.generateTextureReduced(10.0, [64, 48]), g["gen_tex_sum"], rtol=1e-5, atol=1e-6)
    # rays stay optimisable: gradient reaches _rays through projectRaysToNDC + the fused splat
    laser._rays.requires_grad_(True)
    pts01 = laser.projectRaysToNDC()[:, 0:2] * 0.5 + 0.5
    s, _ = ff.graphics.rasterization.splat_reduce(pts01, 10.0, [64, 48], reduce=("sum",))
    s.sum().backward()
    assert laser._rays.grad is not None and torch.isfinite(laser._rays.grad).all() and laser._rays.grad.abs().max() > 0


def test_scene_facade_with_fake_mitsuba(ff):
    params = fm.demo_params()
    sc = ff.Scene(params)
    assert [m.name() for m in sc.meshes()] == ["mesh-A", "mesh-B"]
    assert sc._camera.name() == "PerspectiveCamera" and sc._projector.name() == "Projector"
    assert sc.light("emit-Spot") is not None and sc.material("mat-Mucosa") is not None
    assert sc.light("emit-Spot").float_attributes().keys() == {"cutoff_angle"}
    a = sc.mesh("mesh-A")
    v0 = torch.tensor(list(params["mesh-A.vertex_positions"])).reshape(-1, 3)
    close(a.get_vertices() + a._centroid_mat[0:3, 3], v0, rtol=1e-6, atol=1e-6)
    a.rotate_z(-np.pi, np.pi)
    sc.mesh("mesh-B").setParent(a)
    sc.mesh("mesh-B").set_randomizable(True)
    sc._camera.translate_x(-0.15, 0.15)
    sc.light("emit-Spot").add_vec3_sampler("intensity.value", ff.sampling.UniformScalarToVec3Sampler(0.1, 10.0))
    sc.material("mat-Mucosa").add_float_key("brdf_0.roughness.value", 0.0, 1.0)
    sc.train()
    torch.manual_seed(5)
    sc.randomize()
    assert params.n_updates == 1
    # oracle replay of the same torch-CUDA draws: mesh-A T,R,S then mesh-B T,R,S, light T,R,+attrs, material, camera T,R
    torch.manual_seed(5)
    dr = lambda n=3: torch.rand(n, device="cuda").cpu()  # noqa: E731
    uT, uR, uS = dr(), dr(), dr()
    rA = O.uniform_between(torch.tensor([0.0, 0.0, -np.pi]), torch.tensor([0.0, 0.0, np.pi]), uR)
    WA = O.compose_world([0, 0, 0], rA, [1, 1, 1], a._centroid_mat[0:3, 3].cpu(), torch.eye(4), True)
    out = torch.tensor(list(params["mesh-A.vertex_positions"])).reshape(-1, 3)
    close(out, O.transform_points(a.get_vertices().cpu(), WA), rtol=1e-5, atol=1e-5)
    b = sc.mesh("mesh-B")
    WB = WA @ O.compose_world([0, 0, 0], [0, 0, 0], [1, 1, 1], b._centroid_mat[0:3, 3].cpu(), torch.eye(4), True)
    outB = torch.tensor(list(params["mesh-B.vertex_positions"])).reshape(-1, 3)
    close(outB, O.transform_points(b.get_vertices().cpu(), WB), rtol=1e-5, atol=1e-5)
    inten = params["emit-Spot.intensity.value"]
    assert len(inten) == 3 and inten[0] == inten[1] == inten[2] and 0.1 <= inten[0] <= 10.0
    assert 0.0 <= params["mat-Mucosa.brdf_0.roughness.value"] <= 1.0
    assert isinstance(params["PerspectiveCamera.to_world"], fm.Transform4f)
    # eval mode is deterministic and repeatable
    sc.eval()
    sc.randomize()
    v1 = list(params["mesh-A.vertex_positions"])
    sc2 = ff.Scene(fm.demo_params())
    sc2.mesh("mesh-A").rotate_z(-np.pi, np.pi)
    sc2.eval()
    sc2.randomize()
    assert v1 == list(sc2._mitsuba_params["mesh-A.vertex_positions"])


def test_batched_randomisation_is_split_invariant(ff):
    """Counter-based RNG: sample i is the same whether drawn in one batch of 64, in 4 batches of 16, or at an
    offset (what makes data-parallel sharding bit-reproducible)."""
    def make():
        sc = ff.Scene(fm.demo_params(seed=3, n_a=1000, n_b=10))
        a = sc.mesh("mesh-A")
        a.rotate(torch.tensor([-1.0, -1.0, -1.0]), torch.tensor([1.0, 1.0, 1.0]))
        a.translate_x(-0.5, 0.5)
        a.scale_y(0.5, 2.0)
        sc._camera.rotate_y(-0.5, 0.5)
        sc.train()
        return sc.batch(seed=42)
    full = make().randomize(64, sample0=0)
    sb = make()
    parts = [sb.randomize(16, sample0=16 * i) for i in range(4)]
    assert torch.equal(full.world, torch.cat([p.world for p in parts]))
    assert torch.equal(full.vertices, torch.cat([p.vertices for p in parts]))
    assert not torch.equal(full.world[0], full.world[1])
    other = make()
    other.seed = 43
    assert not torch.equal(other.randomize(4, sample0=0).world, full.world[:4])
    # statistics of the native stream: uniform on the configured range
    big = make().randomize(4096, sample0=1000)
    tx = big.sampled[:, 0, 0]
    assert tx.min() >= -0.5 and tx.max() <= 0.5 and abs(tx.mean().item()) < 0.02 and abs(tx.var().item() - 1 / 12) < 0.01
    # oracle check of sample 7 from its own sampled values
    s = full.sampled[7].cpu()
    W = O.compose_world(s[0], s[1], s[2], sb.entities[0]._centroid_mat[0:3, 3].cpu(), torch.eye(4), True)
    close(full.world[7, 0], W, rtol=1e-5, atol=1e-5)


def test_animation_gather(ff):
    g = torch.Generator().manual_seed(4)
    V, F = 50, 6
    frames_tr, frames_ev = torch.rand(F, V, 3, generator=g), torch.rand(F, V, 3, generator=g)
    sc = ff.Scene(fm.FakeParams())
    m = ff.entity.Mesh("mesh-V", frames_tr[0].cuda())
    m.add_train_animation(frames_tr.cuda())
    m.add_eval_animation(frames_ev.cuda(), max=F - 1)
    m.scale_x(0.5, 2.0)
    sc._meshes.append(m)
    sc.eval()
    res = sc.batch().randomize(8)
    seq = [0, 1, 2, 3, 4, 5, 0, 1]                      # eval walk, max inclusive (sampling/animation.py:27-34)
    for b in range(8):
        W = res.world[b, 0].cpu()
        close(res.mesh_vertices("mesh-V")[b], O.transform_points(frames_ev[seq[b]], W), rtol=1e-5, atol=1e-5)
    sc.train()
    res = sc.batch(seed=9).randomize(64)
    # every train sample must equal SOME train frame transformed by its world (indices are Philox-drawn)
    used = set()
    for b in range(64):
        W = res.world[b, 0].cpu()
        outs = torch.stack([O.transform_points(frames_tr[f], W) for f in range(F)])
        d = (outs - res.mesh_vertices("mesh-V")[b].cpu()).abs().amax(dim=(1, 2))
        assert d.min() < 1e-4
        used.add(int(d.argmin()))
    assert len(used) == F


def test_laser_out_of_bounds_respawn(ff, golden):
    """projection/laser.py:208-249 through ffb_respawn_rays, with the reference's own torch.rand rows injected."""
    g = golden("laser")
    Laser = ff.projection.Laser
    tr = ff.entity.Transformable("projector")
    K01 = T(g["K01"]).cuda()
    laser = Laser(tr, T(g["respawn_rays"]).cuda(), K01, 60.0, 0.01, 1000.0)
    laser.randomize_laser_out_of_bounds(variates=T(g["respawn_variates"]).cuda())
    assert int(laser.last_respawned) == int(g["respawn_k"]) and int(g["respawn_k"]) > 0
    close(laser._rays, g["respawn_laser"], rtol=1e-5, atol=1e-6)
    laser = Laser(tr, T(g["wide"]).cuda(), K01, 60.0, 0.01, 1000.0)
    laser.randomize_camera_out_of_bounds(T(g["respawn_cam_ndc"]).cuda(), variates=T(g["respawn_cam_variates"]).cuda())
    assert int(laser.last_respawned) == int(g["respawn_cam_k"]) and int(g["respawn_cam_k"]) > 0
    close(laser._rays, g["respawn_cam"], rtol=1e-5, atol=1e-6)
    inside = T(g["respawn_inside"]).cuda()
    laser = Laser(tr, inside.clone(), K01, 60.0, 0.01, 1000.0)
    laser.randomize_laser_out_of_bounds()
    assert int(laser.last_respawned) == 0 and torch.equal(laser._rays, inside)      # untouched, not renormalised
    # device Philox stream: respawned rays land inside the field of view, every ray has unit length, same seed -> same rays
    torch.manual_seed(5)
    a = Laser(tr, T(g["respawn_rays"]).cuda(), K01, 60.0, 0.01, 1000.0)
    a.randomize_laser_out_of_bounds()
    torch.manual_seed(5)
    b = Laser(tr, T(g["respawn_rays"]).cuda(), K01, 60.0, 0.01, 1000.0)
    b.randomize_laser_out_of_bounds()
    assert torch.equal(a._rays, b._rays)
    close(torch.linalg.norm(a._rays, dim=1), torch.ones(a._rays.shape[0]), rtol=1e-6, atol=1e-6)
    ndc = ff.utils.math.transform_points(a._rays, K01)[:, 0:2]
    flipped = ndc.clone(); flipped[:, 1] = 1.0 - flipped[:, 1]      # respawn un-projects through (K @ FLIP_Y)^-1
    assert int(((ndc[:, 0] >= 1.0) | (ndc[:, 0] <= 0.0)).sum()) == 0


def test_batched_hand_off_to_the_parameter_map(ff):
    """BatchResult.write_sample writes sample b like Scene.randomize()'s update_* tail, from one
    host copy of all matrices / attributes and device slices of the vertex buffer."""
    params = fm.demo_params()
    sc = ff.Scene(params)
    a = sc.mesh("mesh-A")
    a.rotate_z(-np.pi, np.pi)
    a.translate_x(-0.2, 0.2)
    sc._camera.translate_x(-0.15, 0.15)
    sc.light("emit-Spot").add_vec3_sampler("intensity.value", ff.sampling.UniformScalarToVec3Sampler(0.1, 10.0))
    sc.material("mat-Mucosa").add_float_key("brdf_0.roughness.value", 0.0, 1.0)
    sc.train()
    res = sc.batch(seed=9).randomize(6, sample0=0)
    world, sampled = res.to_host()
    assert world.shape == tuple(res.world.shape) and np.array_equal(world, res.world.cpu().numpy())
    assert np.array_equal(sampled, res.sampled.cpu().numpy())
    assert res.to_host()[0] is world                                   # cached: one copy per result
    for b in (0, 5):
        before = params.n_updates
        res.write_sample(b)
        assert params.n_updates == before + 1
        out = torch.tensor(list(params["mesh-A.vertex_positions"])).reshape(-1, 3)
        assert torch.equal(out, res.mesh_vertices("mesh-A")[b].cpu())
        cam = params["PerspectiveCamera.to_world"].matrix.torch()[0]
        assert torch.equal(cam, res.entity_world("PerspectiveCamera")[b].cpu())
        inten = params["emit-Spot.intensity.value"]
        assert list(inten) == res.attribute("emit-Spot", "intensity.value")[b].cpu().tolist()
        rough = params["mat-Mucosa.brdf_0.roughness.value"]
        assert float(rough) == float(res.attribute("mat-Mucosa", "brdf_0.roughness.value")[b, 0])


def test_ray_generators(ff, golden):
    """Laser.generate_uniform_rays_by_count (laser.py:40-66) against reference-generated values; generate_random_rays and
    initRandomRays (laser.py:69-92,185-194) draw from torch's CUDA generator (the reference's CPU fixture cannot pin those):
    structure only -- unit length, spawn box, determinism under a seed."""
    g = golden("laser")
    Laser = ff.projection.Laser
    K01, K = T(g["K01"]).cuda(), T(g["K"]).cuda()
    close(Laser.generate_uniform_rays_by_count(5, 4, K01), g["by_count_5x4"], rtol=1e-5, atol=1e-6)
    close(Laser.generate_uniform_rays_by_count(3, 3, K), g["by_count_3x3_K"], rtol=1e-5, atol=1e-6)
    torch.manual_seed(3)
    a = Laser.generate_random_rays(64, K01)
    torch.manual_seed(3)
    b = Laser.generate_random_rays(64, K01)
    assert torch.equal(a, b) and a.shape == (64, 3)
    close(torch.linalg.norm(a, dim=1), torch.ones(64), rtol=1e-6, atol=1e-6)
    back = a.clone(); back[:, 2] *= -1.0                       # undo the final z flip, project: spawned in 0.5 +- 0.05
    ndc = ff.utils.math.transform_points(back, K01)[:, 0:2]
    assert float((ndc - 0.5).abs().max()) <= 0.05 + 1e-4
    laser = Laser(ff.entity.Transformable("projector"), a.clone(), K01, 60.0, 0.01, 1000.0)
    laser.initRandomRays()
    close(torch.linalg.norm(laser._rays, dim=1), torch.ones(64), rtol=1e-6, atol=1e-6)
