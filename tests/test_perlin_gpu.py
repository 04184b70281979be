"""GPU parity of the Perlin material textures against reference-generated fixtures
(tests/golden/perlin.npz): the lattice draws the reference consumed are injected, so the comparison is value for value.
Tolerance: 1e-5 relative plus 1e-5 of the largest magnitude -- a noise value is a cancelling sum of O(1) gradient dot
products (device sincosf / powf vs the CPU's vectorised versions differ by an ulp or two on those terms)."""
import random

import numpy as np
import pytest
import torch

from oracle import ff_oracle as O
from test_oracle_golden import _perlin_angles

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def S():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import fireflies_b200.sampling.noise_texture_lerp as S
    return S


def close(a, b, rtol=1e-5, atol=2e-6):
    a = a.detach().cpu().double().numpy() if torch.is_tensor(a) else np.asarray(a, np.float64)
    b = b.detach().cpu().double().numpy() if torch.is_tensor(b) else np.asarray(b, np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    err = np.abs(a - b) - (atol + rtol * np.abs(b))
    assert err.max() <= 0, f"max violation {err.max():.3e}; max abs diff {np.abs(a - b).max():.3e}"


@pytest.mark.parametrize("name", ["a", "b", "c"])
def test_perlin_noise_fixture(S, golden, name):
    g = golden("perlin")
    shape, res, oc, pers, angles = _perlin_angles(g, name)
    noise, tex = S.perlin_texture(shape, res, oc, pers, angles)
    assert tex is None
    close(noise, g[f"{name}_noise"], rtol=1e-5, atol=1e-5 * float(np.abs(g[f"{name}_noise"]).max()))
    ca, cb = torch.tensor([0.2, 0.5, 0.9]), torch.tensor([1.0, 0.0, 0.25])
    _, tex = S.perlin_texture(shape, res, oc, pers, angles, ca.cuda(), cb.cuda())
    close(tex, O.noise_texture_lerp(torch.from_numpy(g[f"{name}_noise"]), ca, cb), rtol=1e-5, atol=2e-5)


def test_noise_texture_sampler_consumes_the_reference_streams(S, golden):
    g = golden("perlin")
    ca, cb = torch.from_numpy(g["sampler_ca"]).cuda(), torch.from_numpy(g["sampler_cb"]).cuda()
    smp = S.NoiseTextureLerpSampler(ca, cb, [128, 128])
    random.seed(31); torch.manual_seed(31)
    tex = smp.sample_train()
    assert tex.shape == (3, 128, 128) and tex.is_cuda
    close(tex, g["sampler_tex"], rtol=1e-5, atol=2e-5)
    assert smp.sample_eval().shape == (3, 128, 128)


def test_perlin_rejects_shapes_the_reference_cannot_tile(S):
    with pytest.raises(RuntimeError):
        S.perlin_texture([100, 100], (8, 8), 2, 0.5, S.perlin_angles((8, 8), 2))
