"""us_quadrature_rules: the 1-D tables of roots() built on the device, against
the reference's host call scipy.special.roots_jacobi, against 50-digit arithmetic, and through
integrate() against the reference's golden integrals."""
import ctypes as C
import os

import mpmath as mp
import numpy as np
import pytest
import torch
from scipy.special import roots_jacobi

import ultrasphere_b200 as us
from ultrasphere_b200 import _native
from tests.helpers import GOLDEN, INTEGRAL_REL, make

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
QD = np.load(os.path.join(GOLDEN, "quadrature.npz"))
QUAD_SPECS = sorted({k.split("|")[0] for k in QD.files})
KIND = {"a": 0, "b": 1, "b'": 2, "c": 3}


def device_rule(kind, n, alpha, beta, dtype=torch.float64):
    """One axis through the C ABI; returns (x, w) as numpy arrays."""
    count = 2 * n if kind == "a" else n
    x = torch.empty(count, dtype=dtype, device=DEV)
    w = torch.empty(count, dtype=dtype, device=DEV)
    rc = _native.lib().us_quadrature_rules(
        1,
        (C.c_uint8 * 1)(KIND[kind]),
        (C.c_int64 * 1)(count),
        (C.c_double * 1)(alpha),
        (C.c_double * 1)(beta),
        (C.c_int32 * 1)(0 if dtype == torch.float32 else 1),
        (C.c_void_p * 1)(x.data_ptr()),
        (C.c_void_p * 1)(w.data_ptr()),
        C.c_void_p(torch.cuda.current_stream().cuda_stream),
    )
    _native.check(rc, "us_quadrature_rules")
    return x.cpu().numpy(), w.cpu().numpy()


def reference_rule(kind, n, alpha, beta):
    """What roots() of the reference derives from scipy (_integral.py:88-105)."""
    t, w = roots_jacobi(n, alpha, beta)
    if kind == "b":
        return np.acos(t), w
    if kind == "b'":
        return np.asin(t), w
    return np.acos(t) / 2, w / 2 ** (alpha + beta + 2)


PARAMS = [(0.0, 0.0), (0.5, 0.5), (1.0, 1.0), (3.5, 3.5), (16.0, 16.0), (0.0, 0.5), (2.0, 0.5), (1.5, 7.0), (15.5, 0.0), (0.0, 15.5), (0.3, 1.7)]


@pytest.mark.parametrize("n", [1, 2, 3, 5, 8, 16, 33, 96, 192, 257, 512])
def test_gauss_tables_against_scipy(n):
    for alpha, beta in PARAMS:
        kinds = ("b", "b'", "c") if alpha == beta else ("c",)
        for kind in kinds:
            x, w = device_rule(kind, n, alpha, beta)
            xr, wr = reference_rule(kind, n, alpha, beta)
            # angles: both sides carry |dt| <= 1 ulp through acos/asin, whose slope is 1/sqrt(1-t^2)
            ts = roots_jacobi(n, alpha, beta)[0]
            slope = 1.0 / np.sqrt(np.maximum(1 - ts * ts, 1e-300))
            assert np.all(np.abs(x - xr) <= 4.5e-16 * slope + 9e-16), (kind, n, alpha, beta)
            # weights: scipy's own error reaches 4e-11 of the largest weight (2e-9 relative on the
            # tiny end weights) at n = 512; test_gauss_tables_against_50_digit_arithmetic is the
            # tight check
            assert np.abs(w - wr).max() <= (2e-14 if n <= 8 else 2e-10) * wr.max(), (kind, n, alpha, beta)
            assert np.abs(w / wr - 1).max() <= 1e-8, (kind, n, alpha, beta)
            assert abs(w.sum() / wr.sum() - 1) <= 1e-14
            assert np.all(np.diff(x) < 0) if kind in ("b", "c") else np.all(np.diff(x) > 0)


def exact_rule(n, alpha, beta, t0):
    """Gauss-Jacobi nodes/weights in 50-digit arithmetic (Newton from approximate nodes t0)."""
    mp.mp.dps = 50
    a, b = mp.mpf(alpha), mp.mpf(beta)
    s = a + b
    diag = [(b - a) / (2 + s)] + [(b * b - a * a) / ((2 * k + s) * (2 * k + s + 2)) for k in range(1, n)]
    off = [mp.mpf(0)] + [
        2 / (2 * k + s) * mp.sqrt(k * (k + a) * (k + b) * (k + s) / ((2 * k + s + 1) * (2 * k + s - 1)))
        if k > 1
        else 2 / (2 + s) * mp.sqrt((1 + a) * (1 + b) / (3 + s))
        for k in range(1, n)
    ]
    mu0 = 2 ** (s + 1) * mp.beta(a + 1, b + 1)
    ts, ws = [], []
    for t in t0:
        t = mp.mpf(float(t))
        for it in range(6):
            pm, p, dm, d, ssq = mp.mpf(0), mp.mpf(1), mp.mpf(0), mp.mpf(0), mp.mpf(1)
            for k in range(n - 1):
                pn = ((t - diag[k]) * p - off[k] * pm) / off[k + 1]
                dn = (p + (t - diag[k]) * d - off[k] * dm) / off[k + 1]
                ssq += pn * pn
                pm, p, dm, d = p, pn, d, dn
            pn = (t - diag[n - 1]) * p - off[n - 1] * pm
            dn = p + (t - diag[n - 1]) * d - off[n - 1] * dm
            if it < 5:
                t = t - pn / dn
        ts.append(t)
        ws.append(mu0 / ssq)
    return ts, ws


@pytest.mark.parametrize("n, alpha, beta", [(7, 0.0, 0.0), (24, 0.5, 0.5), (40, 2.0, 0.5), (64, 15.5, 0.0), (96, 1.0, 1.0), (128, 0.3, 1.7)])
def test_gauss_tables_against_50_digit_arithmetic(n, alpha, beta):
    """Nodes within 1 ulp, weights within 1e-12 relative of the exact rule — tighter than the host
    rule the reference uses, which is the only other authority for these tables."""
    ts, ws = roots_jacobi(n, alpha, beta)
    te, we = exact_rule(n, alpha, beta, ts)
    theta, w = device_rule("c", n, alpha, beta)
    scale = 2 ** (alpha + beta + 2)
    err_t = max(abs(float(mp.mpf(float(theta[i])) * 2 - mp.acos(te[i]))) * float(mp.sqrt(1 - te[i] ** 2)) for i in range(n))
    err_w = max(abs(float(mp.mpf(float(w[i])) * scale / we[i] - 1)) for i in range(n))
    err_w_scipy = max(abs(float(mp.mpf(float(ws[i])) / we[i] - 1)) for i in range(n))
    assert err_t <= 5e-16, err_t
    assert err_w <= 1e-12, (err_w, err_w_scipy)


@pytest.mark.parametrize("tdt", [torch.float32, torch.float64])
@pytest.mark.parametrize("n", [1, 3, 96, 1000, 70000])
def test_trapezoid_axis_is_bit_identical_to_the_reference_expression(n, tdt):
    x, w = device_rule("a", n, 0.0, 0.0, tdt)
    xr = (torch.arange(2 * n, device=DEV, dtype=tdt) * torch.pi / n).cpu().numpy()
    wr = (torch.ones(2 * n, device=DEV, dtype=tdt) * torch.pi / n).cpu().numpy()
    assert x.dtype == xr.dtype and np.array_equal(x, xr) and np.array_equal(w, wr)
    # and within one ulp of the correctly rounded NumPy expression the oracle evaluates
    xn = (np.arange(2 * n, dtype=x.dtype) * x.dtype.type(np.pi) / x.dtype.type(n)).astype(x.dtype)
    assert np.all(np.abs(x - xn) <= np.spacing(np.abs(xn)))


def test_symmetric_rules_mirror_exactly_and_integrate_polynomials():
    for n, lam in ((5, 0.0), (8, 1.5), (33, 4.0)):
        th, w = device_rule("b", n, lam, lam)
        t = np.cos(th)
        assert np.array_equal(w, w[::-1])
        assert np.abs(t + t[::-1]).max() <= 5e-16
        if n % 2:
            assert th[n // 2] == np.acos(0.0)
        # moments of (1-t^2)^lam up to degree 2n-1 against scipy's rule
        ts, ws = roots_jacobi(n, lam, lam)
        for k in range(0, 2 * n, 2):
            assert abs((w * t**k).sum() - (ws * ts**k).sum()) <= 1e-13 * ws.sum()


@pytest.mark.parametrize("spec", QUAD_SPECS)
@pytest.mark.parametrize("dtype", [torch.float64, torch.float32, None])
def test_roots_device_backend_matches_scipy_backend(spec, dtype):
    c = make(us, spec)
    for n in (4, 7, 12):
        xs, ws = us.roots(c, n, expand_dims_x=False, xp=torch, device=DEV, dtype=dtype)
        with us.rule_backend("device"):
            xd, wd = us.roots(c, n, expand_dims_x=False, xp=torch, device=DEV, dtype=dtype)
            xe, we = us.roots(c, n, expand_dims_x=True, expand_dims_w=True, xp=torch, device=DEV, dtype=dtype)
        assert list(xd) == list(xs) == list(c.s_nodes)
        for i, node in enumerate(c.s_nodes):
            assert xd[node].dtype == xs[node].dtype and wd[node].dtype == ws[node].dtype, (node, dtype)
            assert xd[node].shape == xs[node].shape and xd[node].device == xs[node].device
            tol = 2e-14 if xs[node].dtype == torch.float64 else 3e-7
            assert (xd[node] - xs[node]).abs().max().item() <= (8e-15 if xs[node].dtype == torch.float64 else 1e-6)
            assert ((wd[node] - ws[node]).abs() / ws[node].abs().max()).max().item() <= tol
            shp = [1] * c.s_ndim
            shp[i] = xs[node].numel()
            assert list(xe[node].shape) == shp and list(we[node].shape) == shp


@pytest.mark.parametrize("spec", QUAD_SPECS)
def test_integrals_with_device_rules_match_reference_golden(spec):
    c = make(us, spec)
    k = torch.from_numpy(QD[f"{spec}|k"]).to(DEV)

    def f(s):
        return torch.exp(torch.einsum("v,v...->...", k, c.to_cartesian(s, as_array=True)))

    def fsep(s):
        return {node: torch.cos(s[node]) ** 2 for node in c.s_nodes}

    for n in (4, 7, 12):
        if f"{spec}|n{n}|dense" not in QD.files:
            continue
        want = float(QD[f"{spec}|n{n}|dense"])
        with us.rule_backend("device"):
            got = us.integrate(c, f, False, n, xp=torch, device=DEV, dtype=torch.float64)
            sep = us.integrate(c, fsep, True, n, xp=torch, device=DEV, dtype=torch.float64)
        host_sep = us.integrate(c, fsep, True, n, xp=torch, device=DEV, dtype=torch.float64)
        assert abs(got.item() - want) <= INTEGRAL_REL * abs(want), (spec, n)
        for node in c.s_nodes:
            assert abs(sep[node].item() - host_sep[node].item()) <= INTEGRAL_REL * max(1.0, abs(host_sep[node].item()))


def test_full_size_quadrature_with_device_rules_matches_analytic():
    """C5 (create_standard(4), n = 96) with tables from the device: analytic value to 1e-12."""
    from scipy.special import iv

    c = us.create_standard(4)
    k = torch.tensor(np.random.default_rng(0).uniform(0, 1, 5), device=DEV)
    kn = float(torch.linalg.vector_norm(k))
    want = (2 * np.pi) ** 2.5 * kn ** (1 - 2.5) * iv(1.5, kn)
    with us.rule_backend("device"):
        got = us.integrate(
            c, lambda s: torch.exp(torch.einsum("v,v...->...", k, c.to_cartesian(s, as_array=True))), False, 96,
            xp=torch, device=DEV, dtype=torch.float64,
        )
    assert abs(got.item() - want) <= INTEGRAL_REL * want


def test_capacity_arguments_and_many_axes():
    with pytest.raises(ValueError):
        device_rule("b", _native.US_MAX_RULE_POINTS + 1, 0.0, 0.0)
    x, w = device_rule("b", _native.US_MAX_RULE_POINTS, 0.5, 0.5)
    xr, wr = reference_rule("b", _native.US_MAX_RULE_POINTS, 0.5, 0.5)
    assert np.abs(np.cos(x) - np.cos(xr)).max() <= 4.5e-16 and abs(w.sum() / wr.sum() - 1) <= 1e-13
    with pytest.raises(RuntimeError):
        device_rule("b", 4, -1.0, 0.0)
    with pytest.raises(ValueError), us.rule_backend("device"):
        us.roots(us.create_spherical(), _native.US_MAX_RULE_POINTS + 1, expand_dims_x=False, xp=torch, device=DEV)
    # a tree with more angles than one launch holds (US_MAX_AXES): two launches, same tables
    c = us.create_standard(_native.US_MAX_AXES + 6)
    xs, ws = us.roots(c, 6, expand_dims_x=False, xp=torch, device=DEV, dtype=torch.float64)
    with us.rule_backend("device"):
        xd, wd = us.roots(c, 6, expand_dims_x=False, xp=torch, device=DEV, dtype=torch.float64)
    for node in c.s_nodes:
        assert (xd[node] - xs[node]).abs().max().item() <= 8e-15
        assert ((wd[node] - ws[node]).abs() / ws[node].abs().max()).max().item() <= 1e-13
