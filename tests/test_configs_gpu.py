"""GPU parity at the shapes BASELINE.json names besides the bench workload (configs[2]):
configs[1] (vocal-fold scene: 18x18 grid pattern, animated meshes, batch 32), configs[3] (64 x 1024^2 frames, blur + noise)
and the sharding rule of configs[4] (2048 scenes split over ranks) exercised on one GPU."""
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


def close(a, b, rtol=1e-5, atol=1e-6):
    a = a.detach().cpu().double().numpy() if torch.is_tensor(a) else np.asarray(a, np.float64)
    b = b.detach().cpu().double().numpy() if torch.is_tensor(b) else np.asarray(b, np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    err = np.abs(a - b) - (atol + rtol * np.abs(b))
    assert err.max() <= 0, f"max violation {err.max():.3e}; max abs diff {np.abs(a - b).max():.3e}"


def test_config1_vocalfold_scene(ff):
    """main.py:51-77 / examples/vocalfold_scene.py:56-92 without Mitsuba: grid laser -> NDC -> [0,1] -> sum texture at
    500^2, sigma 10 -> (5,5)/(3,3) blur; VocalFold (animated, V=20000, F=64) + Larynx (V=50000) randomised, B=32."""
    from fireflies_b200.postprocessing.base import run_postprocess
    Laser = ff.projection.Laser
    rays = Laser.generate_uniform_rays(0.0275, 18, 18)
    K = ff.utils.io.build_projection_matrix(60, 0.01, 1000.0)
    laser = Laser(ff.entity.Transformable("projector"), rays, K, 60.0, 0.01, 1000.0)
    pts01 = laser.projectRaysToNDC()[:, 0:2] * 0.5 + 0.5
    tex = ff.graphics.rasterization.rasterize_points_baked_sum(pts01, 10.0, [500, 500])
    ref_pts = O.rays_to_ndc(O.uniform_rays(0.0275, 18, 18), O.build_projection_matrix(60, 0.01, 1000.0))[:, 0:2] * 0.5 + 0.5
    ref_tex = O.reduce_sum(O.splat_dense(ref_pts, 10.0, [500, 500]))
    close(tex, ref_tex)
    blurred = run_postprocess(tex.unsqueeze(0), blur=((5, 5), (3.0, 3.0)))[0]
    close(blurred, O.gaussian_blur2d(ref_tex, (5, 5), (3.0, 3.0)))

    g = torch.Generator().manual_seed(2)
    F, V1, V2, B = 64, 20000, 50000, 32
    frames = torch.rand(F, V1, 3, generator=g) * 2 - 1
    larynx = torch.rand(V2, 3, generator=torch.Generator().manual_seed(3)) * 2 - 1
    sc = ff.Scene(fm.FakeParams())
    vf = ff.entity.Mesh("mesh-VocalFold", frames[0].cuda())
    vf.add_train_animation(frames.cuda())
    vf.add_eval_animation(frames.cuda(), max=F - 1)
    vf.scale_x(0.5, 2.0)
    vf.rotate_y(-0.25, 0.25)
    la = ff.entity.Mesh("mesh-Larynx", larynx.cuda())
    la.scale_x(0.8, 1.2)
    la.rotate_y(-0.1, 0.1)
    sc._meshes += [vf, la]
    sc.train()
    res = sc.batch(seed=5).randomize(B)
    assert res.vertices.shape[0] == B
    for b in (0, 7, 31):
        smp = res.sampled[b].cpu()
        for name, verts_of in (("mesh-VocalFold", None), ("mesh-Larynx", larynx)):
            e = res.batch._entity_index[name]
            rows = res.batch._trs_rows[e]
            t, r, s = (smp[i] if i >= 0 else torch.tensor(d) for i, d in zip(rows, ([0., 0, 0], [0., 0, 0], [1., 1, 1])))
            W = O.compose_world(t, r, s, [0, 0, 0], torch.eye(4), True)
            close(res.entity_world(name)[b], W, rtol=1e-5, atol=1e-5)
            got = res.mesh_vertices(name)[b].cpu()
            if verts_of is not None:
                close(got, O.transform_points(verts_of, W), rtol=1e-5, atol=1e-5)
            else:       # the animated mesh: equal to SOME train frame under this sample's world matrix
                d = torch.stack([(O.transform_points(frames[f], W) - got).abs().max() for f in range(F)])
                assert d.min() < 1e-4
    # eval mode: deterministic frame walk 0,1,2,...
    sc.eval()
    res = sc.batch().randomize(3)
    for b in range(3):
        W = res.entity_world("mesh-VocalFold")[b].cpu()
        close(res.mesh_vertices("mesh-VocalFold")[b], O.transform_points(frames[b], W), rtol=1e-5, atol=1e-5)


def test_config3_postprocessing_frames(ff):
    """64 x 1024^2 frames, blur (3,3)/(5,5) gated p=0.5, noise (0, 0.05) gated p=0.5, gates seeded like main.py:138-142."""
    from fireflies_b200.postprocessing.base import run_postprocess
    B, H, W = 64, 1024, 1024
    frames = torch.rand(B, H, W, generator=torch.Generator().manual_seed(5))
    rng = random.Random(6)
    gates = torch.tensor([O.bernoulli_gates([0.5, 0.5], rng) for _ in range(B)], dtype=torch.uint8)
    x = frames.cuda()
    out = run_postprocess(x, blur=((3, 3), (5.0, 5.0)), noise=(0.0, 0.05), gates=gates.cuda(), seed=7, frame0=0)
    assert out.shape == (B, H, W) and float(out.min()) >= 0.0 and float(out.max()) <= 1.0
    for b in range(B):
        gb, gn = bool(gates[b, 0]), bool(gates[b, 1])
        if not gb and not gn:
            assert torch.equal(out[b], x[b])                     # both gates off: the copy PostProcessor makes
    checked = 0
    for b in range(B):                                           # blur only: against the oracle
        if bool(gates[b, 0]) and not bool(gates[b, 1]) and checked < 3:
            close(out[b], O.gaussian_blur2d(frames[b], (3, 3), (5.0, 5.0)))
            checked += 1
    assert checked > 0
    # noise statistics on a noise-only frame (clip bites at the ends of [0,1]: use the interior)
    nb = next(b for b in range(B) if bool(gates[b, 1]) and not bool(gates[b, 0]))
    mid = (x[nb] > 0.3) & (x[nb] < 0.7)
    z = ((out[nb] - x[nb])[mid] / 0.05)
    assert abs(float(z.mean())) < 0.01 and abs(float(z.std()) - 1.0) < 0.01
    # batching / rank independence: frames 32.. processed alone with frame0 = 32 are identical
    out2 = run_postprocess(x[32:], blur=((3, 3), (5.0, 5.0)), noise=(0.0, 0.05), gates=gates[32:].cuda(), seed=7, frame0=32)
    assert torch.equal(out2, out[32:])


def test_config4_sharded_step_equals_whole_batch(ff):
    """Data-parallel rule of configs[4] on one GPU: two 'ranks' of 8 samples vs one batch of 16 -- randomisation
    bit-identical per global sample index, pattern gradients add up."""
    from fireflies_b200.parallel import shard_samples
    N, ts, sigma, Btot = 300, [256, 192], 36.0, 16
    gen = torch.Generator().manual_seed(3)
    pattern = (torch.rand(N, 2, generator=gen) * 0.9 + 0.05).cuda()
    verts = torch.rand(500, 3, generator=gen) * 2 - 1

    def scene():
        sc = ff.Scene(fm.FakeParams())
        m = ff.entity.Mesh("mesh-S", verts.cuda())
        m.rotate_z(-1.0, 1.0)
        m.translate_x(-0.5, 0.5)
        sc._meshes.append(m)
        sc.train()
        return sc

    whole = ff.PatternStep(N, ts, sigma, Btot, scene_batch=scene().batch(seed=42))
    loss_w, dp_w, res_w = whole.forward_backward(pattern, sample0=0)
    dp_parts, loss_parts, verts_parts = [], [], []
    for rank in range(2):
        first, n = shard_samples(Btot, rank, 2)
        part = ff.PatternStep(N, ts, sigma, n, scene_batch=scene().batch(seed=42))
        loss_p, dp_p, res_p = part.forward_backward(pattern, sample0=first)
        dp_parts.append(dp_p); loss_parts.append(loss_p); verts_parts.append(res_p.vertices)
    assert torch.equal(torch.cat(verts_parts), res_w.vertices)           # randomisation independent of the split
    close(torch.cat(loss_parts), loss_w, rtol=1e-6, atol=0)              # per-sample loss: atomic accumulation order varies
    close(dp_parts[0] + dp_parts[1], dp_w, rtol=1e-4, atol=1e-4 * float(dp_w.abs().max()))


def test_pattern_step_rotates_two_result_buffers(ff):
    """PatternStep writes the randomised samples into two alternating BatchResults (no allocation per step): the result of step i is
    intact after step i + 1, reused by step i + 2, and equal to what a fresh-tensor step (rotate_results=False) produces."""
    N, ts, sigma, B = 200, [256, 192], 36.0, 6
    gen = torch.Generator().manual_seed(9)
    pattern = (torch.rand(N, 2, generator=gen) * 0.9 + 0.05).cuda()
    verts = torch.rand(400, 3, generator=gen) * 2 - 1

    def scene():
        sc = ff.Scene(fm.FakeParams())
        m = ff.entity.Mesh("mesh-R", verts.cuda())
        m.rotate_y(-1.0, 1.0)
        m.translate_z(-0.5, 0.5)
        sc._meshes.append(m)
        sc.train()
        return sc

    rot = ff.PatternStep(N, ts, sigma, B, scene_batch=scene().batch(seed=5))
    fresh = ff.PatternStep(N, ts, sigma, B, scene_batch=scene().batch(seed=5), rotate_results=False)
    res, copies = [], []
    for i in range(4):
        _, dp_r, r = rot.forward_backward(pattern, sample0=i * B)
        _, dp_f, f = fresh.forward_backward(pattern, sample0=i * B)
        assert torch.equal(r.vertices, f.vertices) and torch.equal(r.world, f.world) and torch.equal(r.sampled, f.sampled)
        close(dp_r, dp_f, rtol=1e-5, atol=1e-6 * float(dp_f.abs().max()))
        res.append(r); copies.append(r.vertices.clone())
        if i >= 1:
            assert torch.equal(res[i - 1].vertices, copies[i - 1])          # the previous step's samples are still there
    assert res[2] is res[0] and res[3] is res[1] and res[1] is not res[0]
    assert torch.equal(res[0].vertices, copies[2])                          # step 2 reused step 0's buffers
    with pytest.raises(ValueError):
        rot.scene_batch.randomize(B + 1, out=res[0])


def test_shared_pattern_step_equals_per_sample_step(ff):
    """One pattern for all scenes of a step: the folded-gradient path (one forward, upstream gradients summed over the samples,
    one backward) gives the per-sample path's result."""
    N, ts, sigma, B = 500, [256, 256], 49.0, 12
    gen = torch.Generator().manual_seed(17)
    pattern = (torch.rand(N, 2, generator=gen) * 0.9 + 0.05).cuda()
    gS = torch.randn(B, ts[0], ts[1], generator=gen).cuda()
    gO = torch.randn(B, ts[1], ts[0], generator=gen).cuda()
    per = ff.PatternStep(N, ts, sigma, B, per_sample_points=True)
    sh = ff.PatternStep(N, ts, sigma, B, per_sample_points=False)
    _, dp_a, _ = per.forward_backward(pattern, upstream=(gS, gO))
    _, dp_b, _ = sh.forward_backward(pattern, upstream=(gS, gO))
    close(dp_b, dp_a, rtol=1e-4, atol=1e-4 * float(dp_a.abs().max()))
    assert sh.last[0].shape == (B, ts[0], ts[1]) and torch.equal(sh.last[1][3], per.last[1][3])
    la, da, _ = per.forward_backward(pattern)                # L1(softor, sum) loss
    lb, db, _ = sh.forward_backward(pattern)
    close(lb, la, rtol=1e-5, atol=0)
    close(db, da, rtol=1e-4, atol=1e-4 * float(da.abs().max()))
