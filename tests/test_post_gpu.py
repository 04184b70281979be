"""GPU parity: post-processing (blur with reflect border, white noise, gates) vs the CPU oracle.
The blur's third-party arithmetic (kornia 0.7.1) is absent from the reference tree: parity is pinned to the
oracle's restatement only ("parity unpinned" by the reference, see oracle/ff_oracle.py)."""
import random

import numpy as np
import pytest
import torch

from oracle import ff_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def P():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import fireflies_b200.postprocessing as P
    return P


def close(a, b, rtol=1e-5, atol=1e-6):
    a = a.detach().cpu().double().numpy() if torch.is_tensor(a) else np.asarray(a, np.float64)
    b = b.detach().cpu().double().numpy() if torch.is_tensor(b) else np.asarray(b, np.float64)
    assert a.shape == b.shape
    err = np.abs(a - b) - (atol + rtol * np.abs(b))
    assert err.max() <= 0, f"max violation {err.max():.3e}; max abs diff {np.abs(a - b).max():.3e}"


@pytest.mark.parametrize("shape,ks,sg", [
    ((3, 64, 128), (3, 3), (5.0, 5.0)), ((2, 37, 53), (5, 5), (3.0, 3.0)), ((1, 200, 260), (11, 11), (5.0, 5.0)),
    ((2, 33, 132), (5, 3), (2.0, 1.0)), ((1, 9, 8), (15, 7), (4.0, 2.0)), ((1, 1024, 1024), (3, 3), (5.0, 5.0)),
])
def test_blur_vs_oracle(P, shape, ks, sg):
    from fireflies_b200.postprocessing.base import run_postprocess
    x = torch.rand(shape, generator=torch.Generator().manual_seed(1))
    got = run_postprocess(x.cuda(), blur=(ks, sg))
    close(got, O.gaussian_blur2d(x, ks, sg))


def test_white_noise_bit_exact_with_numpy_stream(P, golden):
    g = golden("postprocess")
    wn = P.WhiteNoise(0.02, 0.05, 1.0)
    np.random.seed(21)
    out = wn.post_process(g["img"].copy())
    assert np.array_equal(out, g["wn_out"])


def test_postprocessor_chain_and_gates(P, golden):
    g = golden("postprocess")
    img = g["img"]
    pp = P.PostProcessor([P.GaussianBlur((3, 3), (5, 5), 0.5), P.WhiteNoise(0.0, 0.05, 0.5)])
    random.seed(6)
    np.random.seed(8)
    outs = [pp.post_process(img) for _ in range(6)]
    rng = random.Random(6)
    np.random.seed(8)
    for i in range(6):
        gates = O.bernoulli_gates([0.5, 0.5], rng)
        assert gates == g["gates_seed6"][i].tolist()
        noise = np.random.normal(np.ones_like(img) * 0.0, np.ones_like(img) * 0.05) if gates[1] else None
        ref = O.post_process(img, {"kernel_size": (3, 3), "sigma": (5, 5)}, {}, gates, noise)
        close(outs[i], ref)
    assert np.array_equal(img, g["img"])          # the input is never modified (postprocessor.py:15 copies)


def test_batched_fused_path(P):
    from fireflies_b200.postprocessing.base import run_postprocess
    B, H, W = 6, 96, 160
    x = torch.rand(B, H, W, generator=torch.Generator().manual_seed(2))
    gates = torch.tensor([[1, 1], [1, 0], [0, 1], [0, 0], [1, 1], [0, 1]], dtype=torch.uint8)
    noise = torch.randn(B, H, W, dtype=torch.float64, generator=torch.Generator().manual_seed(3)) * 0.05
    got = run_postprocess(x.cuda(), blur=((3, 3), (5.0, 5.0)), noise=(0.0, 0.05), gates=gates.cuda(), noise_injected=noise.cuda())
    for b in range(B):
        ref = O.post_process(x[b].numpy(), {"kernel_size": (3, 3), "sigma": (5.0, 5.0)}, {}, gates[b].bool().tolist(),
                             noise[b].numpy())
        close(got[b], ref)
    # native Philox noise: statistics, determinism, independence of batching
    pp = P.PostProcessor([P.GaussianBlur((3, 3), (5, 5), 1.0), P.WhiteNoise(0.0, 0.05, 1.0)])
    flat = torch.full((4, 256, 256), 0.5, device="cuda")
    a = pp.post_process_batch(flat, seed=11, frame0=0)
    b2 = pp.post_process_batch(flat, seed=11, frame0=0)
    assert torch.equal(a, b2)
    c = pp.post_process_batch(flat[:2], seed=11, frame0=2)
    assert torch.equal(c, a[2:])
    z = (a - 0.5) / 0.05
    assert abs(z.mean().item()) < 0.01 and abs(z.std().item() - 1) < 0.01
    assert abs((z ** 3).mean().item()) < 0.05 and abs((z ** 4).mean().item() - 3) < 0.1
    assert not torch.equal(a[0], a[1])
    d = pp.post_process_batch(torch.rand(3, 100, 100, device="cuda") * 4 - 1, seed=1)
    assert d.min() >= 0 and d.max() <= 1


def test_silhouette_batch(golden):
    """ApplySilhouette (apply_silhouette.py:17-40): analytic disc -> 11x11 sigma-5 blur -> product, against the oracle;
    the class draws its disc with random.randint exactly like the reference."""
    import random
    from fireflies_b200.postprocessing.apply_silhouette import ApplySilhouette, run_silhouette
    gen = torch.Generator().manual_seed(8)
    frames = torch.rand(3, 512, 384, generator=gen)
    discs = torch.tensor([[150, 250, 200], [100, 300, 170], [380, 20, 60]], dtype=torch.int32)     # the last one crosses two borders
    out = run_silhouette(frames.cuda(), discs.cuda())
    for b in range(3):
        close(out[b], O.silhouette(frames[b], *discs[b].tolist()))
    inplace = frames.cuda().clone()
    run_silhouette(inplace, discs.cuda(), out=inplace)
    assert torch.equal(inplace, out)
    random.seed(4)
    cx, cy, r = random.randint(100, 200), random.randint(200, 300), random.randint(170, 230)
    random.seed(4)
    res = ApplySilhouette().post_process(frames[0].numpy())
    close(torch.from_numpy(res), O.silhouette(frames[0], cx, cy, r))
    # against the reference itself, run with the real cv2.circle under the same seeds (oracle/make_golden.py::post_cases)
    g = golden("postprocess")
    ones = np.ones((512, 448), dtype=np.float32)
    for i, seed in enumerate(g["silhouette_seeds"].tolist()):
        frame = g["silhouette_grad"] if i == 2 else ones
        random.seed(seed)
        close(torch.from_numpy(ApplySilhouette().post_process(frame.copy())), g["silhouette_out"][i], rtol=1e-5, atol=1e-6)
    out2 = run_silhouette(torch.from_numpy(np.stack([ones, ones, g["silhouette_grad"]])).cuda(), torch.from_numpy(g["silhouette_discs"]).cuda())
    close(out2, g["silhouette_out"], rtol=1e-5, atol=1e-6)
