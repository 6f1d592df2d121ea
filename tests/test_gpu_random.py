"""Device-side random_ball: the reference's distributions, a documented
counter-based stream.  Bit parity with NumPy's PCG64 is not possible; what is pinned instead:
  * the Philox4x32-10 implementation against the Random123 known-answer vector (via a Python model
    that the kernel's uniforms are compared with bit-for-bit),
  * the distributions, with Kolmogorov–Smirnov tests against the closed forms the reference's
    construction implies (and against the oracle's NumPy sampler),
  * shapes, dtypes, key sets, ranges and determinism as in the reference (_random.py:96-192)."""
import numpy as np
import pytest
import torch
from scipy import stats

import ultrasphere_b200 as us

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
M32 = 0xFFFFFFFF


def philox4x32_10(counter, key):
    c = list(counter)
    k0, k1 = key
    for _ in range(10):
        p0 = 0xD2511F53 * c[0]
        p1 = 0xCD9E8D57 * c[2]
        c = [((p1 >> 32) ^ c[1] ^ k0) & M32, p1 & M32, ((p0 >> 32) ^ c[3] ^ k1) & M32, p0 & M32]
        k0 = (k0 + 0x9E3779B9) & M32
        k1 = (k1 + 0xBB67AE85) & M32
    return c


def test_philox_model_known_answer():
    # Random123 kat_vectors: philox4x32-10, zero counter and key / all-ones counter and key
    assert philox4x32_10([0, 0, 0, 0], [0, 0]) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert philox4x32_10([M32] * 4, [M32] * 2) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]


def test_kernel_uniforms_are_the_philox_stream():
    """type='spherical' on create_standard(2): theta0 in [0,pi), theta1 in [0,2pi), r in [0,1) come
    from counter (i, 0, block, 2) under key = seed — reproduced here word for word."""

    class FixedSeed:
        def integers(self, lo, hi):
            return 0x0123456789ABCDEF

    c = us.create_standard(2)
    n = 257
    got = us.random_ball(c, shape=(n,), xp=torch, device=DEV, type="spherical", rng=FixedSeed())
    seed = 0x0123456789ABCDEF
    key = [seed & M32, seed >> 32]
    lo, hi = [0.0, 0.0, 0.0], [np.pi, 2 * np.pi, 1.0]
    want = np.empty((3, n))
    for i in range(n):
        for b in range(2):
            w = philox4x32_10([i, 0, b, 2], key)
            u1 = (((w[0] << 32 | w[1]) >> 11) + 1.0) * 2.0**-53
            u2 = ((w[2] << 32 | w[3]) >> 11) * 2.0**-53
            want[2 * b, i] = lo[2 * b] + (1.0 - u1) * (hi[2 * b] - lo[2 * b])
            if 2 * b + 1 < 3:
                want[2 * b + 1, i] = lo[2 * b + 1] + u2 * (hi[2 * b + 1] - lo[2 * b + 1])
    for row, k in enumerate(["theta0", "theta1", "r"]):
        np.testing.assert_allclose(got[k].cpu().numpy(), want[row], rtol=0, atol=4e-16)


@pytest.mark.parametrize("spec_dim", [2, 3, 5, 10, 33])
@pytest.mark.parametrize("surface", [False, True])
def test_uniform_ball_and_sphere_distribution(spec_dim, surface):
    c = us.create_standard(spec_dim - 1)
    n = 400_000
    x = us.random_ball(c, shape=(n,), xp=torch, device=DEV, rng=np.random.default_rng(7), surface=surface)
    assert x.shape == (spec_dim, n) and x.dtype == torch.float64 and x.device.type == "cuda"
    xh = x.cpu().numpy()
    r = np.linalg.norm(xh, axis=0)
    if surface:
        assert np.abs(r - 1).max() < 1e-14
    else:
        assert r.max() <= 1.0
        # uniform in the ball  <=>  r^d uniform on (0,1)
        assert stats.kstest(r**spec_dim, "uniform").pvalue > 1e-4
    # direction uniform on the sphere  <=>  (x_k/r)^2 ~ Beta(1/2, (d-1)/2) for every k; mean zero
    for k in (0, spec_dim - 1):
        u = (xh[k] / r) ** 2
        assert stats.kstest(u, stats.beta(0.5, (spec_dim - 1) / 2).cdf).pvalue > 1e-4
        assert abs(np.mean(xh[k])) < 6 / np.sqrt(n)
    # uncorrelated coordinates
    assert abs(np.mean(xh[0] * xh[1])) < 6 / np.sqrt(n)
    # same two-sample distribution as the reference's NumPy construction (oracle-side restatement)
    rng = np.random.default_rng(11)
    g = rng.normal(0, np.sqrt(0.5), (spec_dim, 100_000))
    ref = g / np.linalg.norm(g, axis=0) if surface else g / np.sqrt(np.sum(g**2, axis=0) + rng.exponential(1, 100_000))
    assert stats.ks_2samp(xh[0, :100_000], ref[0]).pvalue > 1e-4
    assert stats.ks_2samp(np.linalg.norm(xh[:, :100_000], axis=0), np.linalg.norm(ref, axis=0)).pvalue > 1e-4 or surface


def test_spherical_type_ranges_keys_and_determinism():
    c = us.create_from_branching_types("cb'aba")
    s = us.random_ball(c, shape=(3, 50_000), xp=torch, device=DEV, dtype=torch.float32, type="spherical", rng=np.random.default_rng(3))
    assert list(s) == list(c.s_nodes) + ["r"]
    rng_of = {"a": (0, 2 * np.pi), "b": (0, np.pi), "b'": (-np.pi / 2, np.pi / 2), "c": (0, np.pi / 2)}
    for node in c.s_nodes:
        lo, hi = rng_of[c.branching_types[node].value]
        v = s[node].cpu().numpy().astype(np.float64).ravel()
        assert s[node].shape == (3, 50_000) and s[node].dtype == torch.float32
        assert v.min() >= lo - 1e-6 and v.max() <= hi + 1e-6
        assert stats.kstest((v - lo) / (hi - lo), "uniform").pvalue > 1e-4
    assert 0 <= s["r"].min().item() and s["r"].max().item() <= 1
    surf = us.random_ball(c, shape=(10,), xp=torch, device=DEV, type="spherical", surface=True)
    assert "r" not in surf and list(surf) == list(c.s_nodes)
    # same generator state -> same points; a different state -> different points
    a = us.random_ball(c, shape=(1000,), xp=torch, device=DEV, rng=np.random.default_rng(5))
    b = us.random_ball(c, shape=(1000,), xp=torch, device=DEV, rng=np.random.default_rng(5))
    d = us.random_ball(c, shape=(1000,), xp=torch, device=DEV, rng=np.random.default_rng(6))
    assert torch.equal(a, b) and not torch.equal(a, d)
    # a point's value does not depend on the batch it is generated in
    big = us.random_ball(c, shape=(5000,), xp=torch, device=DEV, rng=np.random.default_rng(5))
    assert torch.equal(big[:, :1000], a)
    with pytest.raises(ValueError):
        us.random_ball(c, shape=(2,), xp=torch, device=DEV, type="gaussian")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        us.random_ball(c, shape=(2,), xp=torch, device="cpu")
    assert us.random_ball(c, shape=(0,), xp=torch, device=DEV).shape == (c.c_ndim, 0)


def test_random_points_feed_the_transforms():
    """random_ball -> from_cartesian -> to_cartesian stays on the device and round-trips."""
    c = us.create_hopf(3)
    x = us.random_ball(c, shape=(1 << 20,), xp=torch, device=DEV, dtype=torch.float64, rng=np.random.default_rng(1))
    s = c.from_cartesian(x)
    assert s["r"].max().item() <= 1.0
    back = c.to_cartesian(s, as_array=True)
    assert (back - x).abs().max().item() < 1e-15


@pytest.mark.parametrize("dim", [3, 10, 16, 17, 40])
def test_float32_ball_distribution(dim):
    """The float32 kernel (single-precision Box–Muller, 4 normals per Philox block; dims > 16
    regenerate instead of keeping their normals in registers)."""
    c = us.create_standard(dim - 1)
    n = 300_000
    x = us.random_ball(c, shape=(n,), xp=torch, device=DEV, dtype=torch.float32, rng=np.random.default_rng(9))
    assert x.dtype == torch.float32 and x.shape == (dim, n)
    xh = x.cpu().numpy().astype(np.float64)
    r = np.linalg.norm(xh, axis=0)
    assert r.max() <= 1.0 + 1e-6
    assert stats.kstest(r**dim, "uniform").pvalue > 1e-4
    for k in (0, dim // 2, dim - 1):
        assert stats.kstest((xh[k] / r) ** 2, stats.beta(0.5, (dim - 1) / 2).cdf).pvalue > 1e-4
    s = us.random_ball(c, shape=(n,), xp=torch, device=DEV, dtype=torch.float32, surface=True)
    assert (torch.linalg.vector_norm(s.double(), dim=0) - 1).abs().max().item() < 5e-7
