"""Accuracy of the library's own float32 device math (us_common.cuh) against float64, and the
bit-exactness of its IEEE square root.  These bounds are what the parity tolerances rest on."""
import numpy as np
import pytest
import torch

from ultrasphere_b200 import _native

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def probe(fn, a, b=None):
    at = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(DEV)
    bt = None if b is None else torch.from_numpy(np.ascontiguousarray(b, dtype=np.float32)).to(DEV)
    o0 = torch.empty_like(at)
    o1 = torch.empty_like(at)
    st = torch.cuda.current_stream().cuda_stream
    _native.check(
        _native.lib().us_math_probe_f32(fn, at.numel(), at.data_ptr(), None if bt is None else bt.data_ptr(), o0.data_ptr(), o1.data_ptr(), st),
        "us_math_probe_f32",
    )
    return o0.cpu().numpy(), o1.cpu().numpy()


def ulp_err(got, exact):
    exact32 = exact.astype(np.float32)
    return np.abs(got.astype(np.float64) - exact) / np.spacing(np.abs(exact32)).astype(np.float64)


def test_sincos_accuracy():
    rng = np.random.default_rng(0)
    parts = [rng.uniform(-np.pi, np.pi, 2_000_000), rng.uniform(-100, 100, 2_000_000), rng.uniform(-32768, 32768, 2_000_000),
             rng.uniform(-1e-3, 1e-3, 500_000)]
    k = np.arange(-20000, 20001)
    for d in (0, 1, -1, 3):  # floats next to multiples of pi/2: exact zeros of sin / cos
        x = (k * (np.pi / 2)).astype(np.float32)
        for _ in range(abs(d)):
            x = np.nextafter(x, np.float32(np.inf if d > 0 else -np.inf))
        parts.append(x.astype(np.float64))
    x = np.concatenate(parts).astype(np.float32)
    s, c = probe(0, x)
    xd = x.astype(np.float64)
    assert ulp_err(s, np.sin(xd)).max() <= 2.0
    assert ulp_err(c, np.cos(xd)).max() <= 2.0
    # relative accuracy is kept at the zeros (Cody–Waite with an exact first step)
    assert (np.abs(s - np.sin(xd)) <= 2.5e-7 * np.abs(np.sin(xd))).all()
    assert (np.abs(c - np.cos(xd)) <= 2.5e-7 * np.abs(np.cos(xd))).all()


def test_sincos_slow_path_and_specials():
    x = np.float32([32768.0, -32768.0, 1e6, -3.5e7, 1e20, 3.0e38, np.inf, -np.inf, np.nan, 0.0, -0.0])
    s, c = probe(0, x)
    xd = x.astype(np.float64)
    with np.errstate(invalid="ignore"):
        ws, wc = np.sin(xd), np.cos(xd)
    fin = np.isfinite(xd)
    assert ulp_err(s[fin], ws[fin]).max() <= 2.0 and ulp_err(c[fin], wc[fin]).max() <= 2.0
    assert np.isnan(s[~fin]).all() and np.isnan(c[~fin]).all()


def test_atan2_accuracy_and_conventions():
    rng = np.random.default_rng(1)
    y = np.concatenate([rng.uniform(-1, 1, 3_000_000), rng.uniform(-1, 1, 500_000) * 1e-6, rng.uniform(-1e20, 1e20, 500_000)])
    x = np.concatenate([rng.uniform(-1, 1, 3_000_000), rng.uniform(-1, 1, 500_000), rng.uniform(-1e20, 1e20, 500_000)])
    y, x = y.astype(np.float32), x.astype(np.float32)
    a, _ = probe(1, y, x)
    exact = np.arctan2(y.astype(np.float64), x.astype(np.float64))
    assert ulp_err(a, exact).max() <= 4.5
    assert (np.abs(a - exact) <= 3.5e-7 * np.abs(exact)).all()
    # IEEE conventions (signed zeros, infinities, NaN, denormals, huge) via the libdevice path
    sp = np.float32([0.0, -0.0, 1.0, -1.0, np.inf, -np.inf, np.nan, 1e-40, -1e-40, 3e38, -3e38, 1e-35])
    yy, xx = [g.ravel() for g in np.meshgrid(sp, sp)]
    a, _ = probe(1, yy, xx)
    want = np.arctan2(yy, xx)
    assert np.array_equal(np.isnan(a), np.isnan(want))
    ok = ~np.isnan(want)
    assert np.array_equal(np.signbit(a[ok]), np.signbit(want[ok]))
    assert (np.abs(a[ok].astype(np.float64) - want[ok]) <= 4 * np.spacing(np.abs(want[ok])).astype(np.float64)).all()


def test_sqrt_is_bit_identical_to_ieee():
    rng = np.random.default_rng(2)
    x = np.concatenate([
        rng.uniform(0, 4, 4_000_000), 10.0 ** rng.uniform(-44, 38, 2_000_000), (rng.uniform(0, 1, 1_000_000) ** 2),
        np.float64([0.0, -0.0, np.inf, np.nan, -1.0, 1e-45, 1.1754944e-38, 3.4028235e38, 1.0, 4.0, 2.0]),
    ]).astype(np.float32)
    mine, ieee = probe(2, x)
    assert np.array_equal(mine.view(np.uint32)[~np.isnan(ieee)], ieee.view(np.uint32)[~np.isnan(ieee)])
    assert np.array_equal(np.isnan(mine), np.isnan(ieee))
    with np.errstate(invalid="ignore"):
        want = np.sqrt(x)
    assert np.array_equal(ieee[~np.isnan(want)], want[~np.isnan(want)])  # and both equal NumPy's sqrt


# ---- float64 node math ---------------------------------------------------------------------------
def probe64(fn, a, b=None):
    at = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(DEV)
    bt = None if b is None else torch.from_numpy(np.ascontiguousarray(b, dtype=np.float64)).to(DEV)
    o0 = torch.empty_like(at)
    o1 = torch.empty_like(at)
    st = torch.cuda.current_stream().cuda_stream
    _native.check(
        _native.lib().us_math_probe_f64(fn, at.numel(), at.data_ptr(), None if bt is None else bt.data_ptr(), o0.data_ptr(), o1.data_ptr(), st),
        "us_math_probe_f64",
    )
    return o0.cpu().numpy(), o1.cpu().numpy()


def test_f64_atan2_and_norm_accuracy():
    """The float64 node op against 50-digit mpmath on a sample, and against NumPy on millions."""
    import mpmath as mp

    rng = np.random.default_rng(3)
    n = 4_000_000
    y = np.concatenate([rng.uniform(-1, 1, n), rng.uniform(-1, 1, n // 4) * 1e-9, rng.uniform(-1e100, 1e100, n // 4), rng.uniform(0, 1, n // 4)])
    x = np.concatenate([rng.uniform(-1, 1, n), rng.uniform(-1, 1, n // 4), rng.uniform(-1e100, 1e100, n // 4), rng.uniform(-1, 1, n // 4) * 1e-7])
    a, nrm = probe64(1, y, x)
    want = np.arctan2(y, x)
    d = np.abs(a - want) / np.spacing(np.abs(want))
    assert d.max() <= 2.0, d.max()  # NumPy's own atan2 is within 1 ulp of exact
    assert np.array_equal(nrm, np.sqrt(y * y + x * x))  # IEEE: squares, add, sqrt each rounded once
    mp.mp.dps = 50
    idx = rng.choice(len(y), 3000, replace=False)
    worst = 0.0
    for i in idx:
        exact = mp.atan2(mp.mpf(float(y[i])), mp.mpf(float(x[i])))
        err = abs(mp.mpf(float(a[i])) - exact) / mp.mpf(float(np.spacing(abs(float(exact)))))
        worst = max(worst, float(err))
    assert worst <= 1.5, worst  # libdevice documents 2 ulp for atan2
    # IEEE conventions through the libdevice path
    sp = np.float64([0.0, -0.0, 1.0, -1.0, np.inf, -np.inf, np.nan, 1e-320, -1e-320, 1e308, -1e308, 1e-200, 1e200])
    yy, xx = [g.ravel() for g in np.meshgrid(sp, sp)]
    a, nrm = probe64(1, yy, xx)
    want = np.arctan2(yy, xx)
    assert np.array_equal(np.isnan(a), np.isnan(want))
    ok = ~np.isnan(want)
    assert np.array_equal(np.signbit(a[ok]), np.signbit(want[ok]))
    assert (np.abs(a[ok] - want[ok]) <= 2 * np.spacing(np.abs(want[ok]))).all()
    with np.errstate(all="ignore"):
        wn = np.sqrt(yy * yy + xx * xx)
    assert np.array_equal(nrm, wn, equal_nan=True)


def test_f64_sqrt_is_bit_identical_to_ieee():
    rng = np.random.default_rng(4)
    x = np.concatenate([
        rng.uniform(0, 4, 6_000_000), 10.0 ** rng.uniform(-299, 299, 3_000_000), rng.uniform(0, 1, 1_000_000) ** 2,
        np.nextafter(np.arange(1, 200001, dtype=np.float64) ** 2, 0), np.nextafter(np.arange(1, 200001, dtype=np.float64) ** 2, np.inf),
        np.float64([1.0, 2.0, 4.0, 0.25, 1e-300 * 1.0000001, 9.99e299]),
    ])
    mine, ieee = probe64(2, x)
    assert np.array_equal(mine, ieee)
    assert np.array_equal(ieee, np.sqrt(x))


def test_f64_sincos_accuracy():
    """<= 0.75 ulp against 50-digit mpmath on a sample (libdevice documents 2, NumPy measures ~0.5),
    <= 1 ulp away from NumPy on millions of arguments, relative accuracy kept at the zeros."""
    import mpmath as mp

    rng = np.random.default_rng(5)
    k = np.arange(-10000, 10001, dtype=np.float64)
    near = k * (np.pi / 2)
    x = np.concatenate([
        rng.uniform(-np.pi, np.pi, 3_000_000), rng.uniform(-100, 100, 1_000_000), rng.uniform(-16384, 16384, 1_000_000),
        rng.uniform(-1e-4, 1e-4, 200_000), near, np.nextafter(near, np.inf), np.nextafter(near, -np.inf),
    ])
    s, c = probe64(0, x)
    ws, wc = np.sin(x), np.cos(x)
    assert (np.abs(s - ws) <= 1.0 * np.spacing(np.abs(ws))).all()
    assert (np.abs(c - wc) <= 1.0 * np.spacing(np.abs(wc))).all()
    same = np.mean(s == ws), np.mean(c == wc)
    assert min(same) > 0.9, same  # most values are bit-identical to NumPy's
    mp.mp.dps = 50
    idx = np.concatenate([rng.choice(5_000_000, 2500, replace=False), 5_200_000 + rng.choice(60_000, 1500, replace=False)])
    worst = 0.0
    for i in idx:
        xv = mp.mpf(float(x[i]))
        for got, exact in ((s[i], mp.sin(xv)), (c[i], mp.cos(xv))):
            if exact == 0:
                continue
            err = abs(mp.mpf(float(got)) - exact) / mp.mpf(float(np.spacing(abs(float(exact)))))
            worst = max(worst, float(err))
    assert worst <= 0.75, worst
    # slow path and specials
    xs = np.float64([16384.0, -16384.0, 1e6, 1e9, -3.5e12, 1e22, 1e300, np.inf, -np.inf, np.nan, 0.0, -0.0])
    s, c = probe64(0, xs)
    with np.errstate(invalid="ignore"):
        ws, wc = np.sin(xs), np.cos(xs)
    fin = np.isfinite(xs)
    assert (np.abs(s[fin] - ws[fin]) <= 2 * np.spacing(np.abs(ws[fin]))).all()
    assert (np.abs(c[fin] - wc[fin]) <= 2 * np.spacing(np.abs(wc[fin]))).all()
    assert np.isnan(s[~fin]).all() and np.isnan(c[~fin]).all()
