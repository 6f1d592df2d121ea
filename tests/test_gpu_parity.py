"""Parity of the CUDA path (through the public API → C ABI) with the NumPy oracle and with the
golden vectors of the unmodified reference.  Needs a B200: run with ``-m gpu``.

Tolerances (see tests/helpers.py and DESIGN.md "Tolerances"):
  float64  r bit-identical; angles <= 4 ulp; leaves |dx| <= 4 ulp(r)  [r = norm of the point]
           and element-wise <= 4 + depth ulp (every ancestor contributes one sin/cos whose
           last-bit rounding differs between libdevice and NumPy's SIMD kernels)
  float32  r bit-identical; angles and leaves <= 2e-6 relative element-wise (floor: 2e-6 * r
           absolute for leaves that are exact zeros of sin/cos)
"""
import json
import os

import numpy as np
import pytest
import torch

import ultrasphere_b200 as us
from oracle import vks_oracle as vo
from tests.helpers import F32_REL, F64_ULP, GOLDEN, make
from ultrasphere_b200._plan import plan_for

pytestmark = pytest.mark.gpu

DEV = "cuda:0"
SPECS = [
    "polar", "spherical", "standard:1", "standard:2", "standard:4", "standard:9", "standard:40",
    "standard_prime:3", "standard_prime:8", "hopf:1", "hopf:2", "hopf:3", "hopf:5",
    "random:1:0", "random:5:0", "random:6:3", "random:32:0", "random:32:1", "random:103:7",
    "expr:cbaa", "expr:cbbaa", "expr:cbabba", "expr:cab'a", "expr:bpbpa",
]
DTYPES = [("f32", np.float32, torch.float32), ("f64", np.float64, torch.float64)]


def dev(a: np.ndarray) -> torch.Tensor:
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def host(t: torch.Tensor) -> np.ndarray:
    return t.detach().cpu().numpy()


def depth_of(c) -> np.ndarray:
    ct = plan_for(c)
    return np.asarray([len(a) for a in ct.leaf_ancestors])


def assert_angles_close(got: np.ndarray, want: np.ndarray, what=""):
    assert got.dtype == want.dtype and got.shape == want.shape
    if got.dtype == np.float64:
        d = vo.ulp_distance(got, want)
        assert d.max() <= F64_ULP, f"{what}: {d.max()} ulp"
    else:
        err = np.abs(got.astype(np.float64) - want.astype(np.float64))
        assert (err <= F32_REL * np.abs(want.astype(np.float64))).all(), f"{what}: rel {np.max(err / np.abs(want))}"


def assert_leaves_close(got: np.ndarray, want: np.ndarray, r: np.ndarray, depth: np.ndarray, what=""):
    """got/want: (c_ndim, *shape); r broadcastable to shape; depth: (c_ndim,)."""
    assert got.dtype == want.dtype and got.shape == want.shape
    g, w = got.astype(np.float64), want.astype(np.float64)
    err = np.abs(g - w)
    rr = np.abs(np.broadcast_to(r, got.shape[1:]).astype(np.float64))[None]
    if got.dtype == np.float64:
        assert (err <= F64_ULP * np.spacing(rr)).all(), f"{what}: {np.max(err / np.spacing(rr))} ulp(r)"
        d = vo.ulp_distance(got, want)
        lim = (F64_ULP + depth).reshape((-1,) + (1,) * (got.ndim - 1))
        assert (d <= lim).all(), f"{what}: {d.max()} ulp element-wise"
    else:
        ok = (err <= F32_REL * np.abs(w)) | (err <= F32_REL * np.finfo(np.float32).eps * rr)
        assert ok.all(), f"{what}: rel {np.nanmax(err / np.abs(w))}"


@pytest.mark.parametrize("dname,ndt,tdt", DTYPES)
@pytest.mark.parametrize("spec", SPECS)
def test_transforms_match_oracle(spec, dname, ndt, tdt):
    c, t = make(us, spec), vo.make(spec)
    n = 20000 if c.s_ndim <= 40 else 3000
    x = np.random.default_rng(11).uniform(-1, 1, (c.c_ndim, n)).astype(ndt)
    want = vo.from_cartesian(t, x)
    got = c.from_cartesian(dev(x))  # array input, indexed by leaf id
    assert list(got) == list(want) == ["r"] + list(reversed(c.s_nodes))
    assert all(v.dtype == tdt and v.shape == (n,) for v in got.values())
    assert np.array_equal(host(got["r"]), want["r"]), "r must be bit-identical"
    for k in c.s_nodes:
        assert_angles_close(host(got[k]), want[k], f"{spec} {k}")
    # to_cartesian fed with the ORACLE's angles: isolates this kernel's error
    sph = {k: dev(v) for k, v in want.items()}
    back = c.to_cartesian(sph)
    wantx = vo.to_cartesian(t, want)
    assert list(back) == list(wantx) == list(c.c_nodes)
    assert_leaves_close(
        np.stack([host(back[k]) for k in c.c_nodes]), np.stack([wantx[k] for k in c.c_nodes]), want["r"], depth_of(c), spec
    )
    # mapping input gives the same bits as array input; as_array stacks the same bits
    got2 = c.from_cartesian({k: dev(x[i]) for i, k in enumerate(c.c_nodes)})
    for k in got:
        assert torch.equal(got[k], got2[k])
    arr = c.to_cartesian(sph, as_array=True)
    assert arr.shape == (c.c_ndim, n) and all(torch.equal(arr[i], back[k]) for i, k in enumerate(c.c_nodes))
    # GPU round trip against the input
    rt = host(c.to_cartesian(got, as_array=True)).astype(np.float64)
    tol = 8 * np.finfo(ndt).eps
    assert np.max(np.abs(rt - x) / want["r"][None]) <= tol


TR = np.load(os.path.join(GOLDEN, "transforms.npz"))
GOLDEN_SPECS = sorted({k.split("|")[0] for k in TR.files})


@pytest.mark.parametrize("dname,ndt,tdt", DTYPES)
@pytest.mark.parametrize("spec", GOLDEN_SPECS)
def test_transforms_match_reference_golden(spec, dname, ndt, tdt):
    """Inputs and outputs recorded from the unmodified reference (oracle/gen_golden.py)."""
    c = make(us, spec)
    p = f"{spec}|{dname}|"
    keys = ["r"] + list(reversed(c.s_nodes))
    got = c.from_cartesian(dev(TR[p + "x"]))
    want = TR[p + "sph"]
    assert np.array_equal(host(got["r"]), want[0])
    for i, k in enumerate(keys[1:], start=1):
        assert_angles_close(host(got[k]), want[i], f"{spec} {k}")
    d = depth_of(c)
    back = c.to_cartesian({k: dev(want[i]) for i, k in enumerate(keys)}, as_array=True)
    assert_leaves_close(host(back), TR[p + "cart"], want[0], d, spec)
    free = {k: dev(TR[p + "ang"][i]) for i, k in enumerate(c.s_nodes)}
    assert_leaves_close(host(c.to_cartesian(free, as_array=True)), TR[p + "cart_free_no_r"], np.ones(1, ndt), d, spec)
    free["r"] = dev(TR[p + "r"])
    assert_leaves_close(host(c.to_cartesian(free, as_array=True)), TR[p + "cart_free"], TR[p + "r"], d, spec)


@pytest.mark.parametrize("dname,ndt,tdt", DTYPES)
@pytest.mark.parametrize("spec", GOLDEN_SPECS)
def test_edge_vectors(spec, dname, ndt, tdt):
    """zeros, signed zeros, axis points, 1e±30 (float32 overflows to inf exactly like NumPy)."""
    c = make(us, spec)
    p = f"{spec}|{dname}|"
    keys = ["r"] + list(reversed(c.s_nodes))
    got = c.from_cartesian(dev(TR[p + "edge_x"]))
    want = TR[p + "edge_sph"]
    assert np.array_equal(host(got["r"]), want[0], equal_nan=True)
    for i, k in enumerate(keys[1:], start=1):
        g, w = host(got[k]), want[i]
        assert np.array_equal(np.isnan(g), np.isnan(w))
        # atan2 conventions at ±0 / inf are IEEE in both; values within a few ulp, signs equal
        ok = np.isnan(w) | (np.abs(g.astype(np.float64) - w.astype(np.float64)) <= 4 * np.spacing(np.abs(w)).astype(np.float64))
        assert ok.all(), (spec, k)
        assert np.array_equal(np.signbit(g[~np.isnan(w)]), np.signbit(w[~np.isnan(w)])), (spec, k)


def test_known_answers_from_reference_doctests():
    with open(os.path.join(GOLDEN, "kat.json")) as fh:
        kat = json.load(fh)
    c = us.create_spherical()
    s = c.from_cartesian(torch.tensor([1.0, 2.0, 3.0], dtype=torch.float64, device=DEV))
    assert list(s) == ["r", "phi", "theta"]
    for key, val in kat["spherical_from_123_rounded"].items():
        assert round(s[key].item(), 6) == val
    assert s["r"].item().hex() == kat["spherical_from_123"]["r"]
    back = c.to_cartesian(s)
    assert {k: round(back[k].item(), 5) for k in c.c_nodes} == {0: 1.0, 1: 2.0, 2: 3.0}
    for name, factory, x in (
        ("hopf2", lambda: us.create_hopf(2), kat["hopf2_x"]),
        ("standard3", lambda: us.create_standard(3), kat["hopf2_x"]),
        ("standard_prime2", lambda: us.create_standard_prime(2), kat["standard_prime2_x"]),
    ):
        got = factory().from_cartesian(torch.tensor(x, dtype=torch.float64, device=DEV))
        assert got["r"].item().hex() == kat[name]["r"]
        for key, hexval in kat[name].items():
            want = float.fromhex(hexval)
            assert abs(got[key].item() - want) <= 4 * np.spacing(abs(want)), (name, key)


# ---- the reference's own transform tests, on the CUDA path (tests/test_coordinates_transform.py) ----
@pytest.mark.parametrize("shape", [(1,), (2, 3), (4, 5, 6)])
def test_spherical_to_3d(shape):
    r = torch.rand(shape, device=DEV)
    theta = torch.rand(shape, device=DEV) * torch.pi
    phi = torch.rand(shape, device=DEV) * 2 * torch.pi
    actual = us.create_spherical().to_cartesian({"r": r, "theta": theta, "phi": phi})
    rsin = r * torch.sin(theta)
    for i, value in enumerate((rsin * torch.cos(phi), rsin * torch.sin(phi), r * torch.cos(theta))):
        assert torch.allclose(value, actual[i], rtol=1e-3, atol=1e-3)
    assert set(actual.keys()) == {0, 1, 2}


@pytest.mark.parametrize("shape", [(1,), (2, 3), (4, 5, 6)])
def test_cartesian_to_spherical_3d(shape):
    x, y, z = (torch.rand(shape, device=DEV) * 2 - 1 for _ in range(3))
    actual = us.create_spherical().from_cartesian({0: x, 1: y, 2: z})
    r = torch.sqrt(x * x + y * y + z * z)
    expected = {"r": r, "theta": torch.acos(z / r), "phi": torch.atan2(y, x)}
    for key, value in expected.items():
        assert torch.allclose(value, actual[key], rtol=1e-3, atol=1e-3)
    assert set(actual.keys()) == {"r", "theta", "phi"}


@pytest.mark.parametrize("shape", [(1,), (2, 3), (2, 4, 6)])
@pytest.mark.parametrize("factory", [us.create_spherical, lambda: us.create_hopf(3), lambda: us.create_random(1), lambda: us.create_random(5)])
def test_cartesian_to_spherical_round_trip(factory, shape):
    c = factory()
    x = torch.rand((c.c_ndim, *shape), device=DEV) * 2 - 1
    back = c.to_cartesian(c.from_cartesian(x))
    back = torch.stack([back[key] for key in range(c.c_ndim)])
    assert torch.allclose(x, back, rtol=1e-6, atol=1e-6)


# ---- input generality -------------------------------------------------------------------------
@pytest.mark.parametrize("dname,ndt,tdt", DTYPES)
def test_broadcast_inputs_keep_reference_shapes(dname, ndt, tdt):
    """Inputs of different shapes: every output keeps the broadcast shape of what feeds it."""
    c, t = us.create_from_branching_types("cbaa"), vo.tree_from_branching("cbaa")
    rng = np.random.default_rng(3)
    n = 7
    ang = {}
    for i, node in enumerate(c.s_nodes):
        shp = [1] * c.s_ndim
        shp[i] = n + i
        ang[node] = rng.uniform(-3, 3, shp).astype(ndt)
    want = vo.to_cartesian(t, dict(ang))
    got = c.to_cartesian({k: dev(v) for k, v in ang.items()})
    d = depth_of(c)
    for i, leaf in enumerate(c.c_nodes):
        assert tuple(got[leaf].shape) == want[leaf].shape, leaf
        assert_leaves_close(host(got[leaf])[None], want[leaf][None], np.ones(1, ndt), d[i : i + 1], f"leaf {leaf}")
    full = c.to_cartesian({k: dev(v) for k, v in ang.items()}, as_array=True)
    wantf = vo.to_cartesian(t, dict(ang), as_array=True)
    assert tuple(full.shape) == wantf.shape
    assert_leaves_close(host(full), wantf, np.ones(1, ndt), d)
    # r as python scalar, 0-d tensor and full tensor
    for rv in (2.5, np.asarray(2.5, dtype=ndt), rng.uniform(0.5, 2, (n, 1, 1, 1)).astype(ndt)):
        a = dict(ang)
        a["r"] = rv
        w = vo.to_cartesian(t, dict(a), as_array=True)
        b = {k: dev(v) for k, v in ang.items()}
        b["r"] = rv if isinstance(rv, float) else dev(rv)
        g = c.to_cartesian(b, as_array=True)
        assert tuple(g.shape) == w.shape
        assert_leaves_close(host(g), w.astype(ndt), np.broadcast_to(np.asarray(rv, dtype=ndt), w.shape[1:]), d)
    # from_cartesian with leaves of different shapes
    xs = {leaf: rng.uniform(-1, 1, [n if j == i % 2 else 1 for j in range(2)]).astype(ndt) for i, leaf in enumerate(c.c_nodes)}
    wants = vo.from_cartesian(t, xs)
    gots = c.from_cartesian({k: dev(v) for k, v in xs.items()})
    assert list(gots) == list(wants)
    for k in wants:
        assert tuple(gots[k].shape) == wants[k].shape, k
        if k == "r":
            assert np.array_equal(host(gots[k]), wants[k])
        else:
            assert_angles_close(host(gots[k]), wants[k], str(k))


def test_strided_expanded_zero_dim_and_empty_inputs():
    c, t = us.create_standard(3), vo.create_standard(3)
    rng = np.random.default_rng(9)
    big = rng.uniform(-1, 1, (4, 6, 10))
    xt = dev(big)
    view = xt[:, ::2, 1::3]  # non-contiguous rows
    got = c.from_cartesian(view)
    want = vo.from_cartesian(t, big[:, ::2, 1::3])
    assert np.array_equal(host(got["r"]), want["r"])
    for k in c.s_nodes:
        assert_angles_close(host(got[k]), want[k])
    # transposed leaves
    tr = {i: xt[i].t() for i in range(4)}
    got = c.from_cartesian(tr)
    want = vo.from_cartesian(t, {i: big[i].T for i in range(4)})
    assert np.array_equal(host(got["r"]), want["r"]) and got["r"].shape == (10, 6)
    # expanded (stride 0) angle
    th = {k: dev(rng.uniform(0, 3, (5,))) for k in c.s_nodes}
    th["theta1"] = th["theta1"][:1].expand(5)
    g = c.to_cartesian(th, as_array=True)
    w = vo.to_cartesian(t, {k: host(v) for k, v in th.items()}, as_array=True)
    assert_leaves_close(host(g), w, np.ones(1), depth_of(c))
    # 0-d tensors
    z = c.from_cartesian({i: torch.tensor(float(big[i, 0, 0]), dtype=torch.float64, device=DEV) for i in range(4)})
    wz = vo.from_cartesian(t, {i: np.asarray(big[i, 0, 0]) for i in range(4)})
    assert z["r"].shape == () and abs(z["r"].item() - float(wz["r"])) <= np.spacing(float(wz["r"]))
    # empty batch
    e = c.from_cartesian(torch.empty((4, 0), device=DEV))
    assert all(v.shape == (0,) for v in e.values())
    assert c.to_cartesian(e, as_array=True).shape == (4, 0)
    # mixed dtypes promote like NumPy / torch
    m = c.to_cartesian({"theta0": dev(np.float32([0.3])), "theta1": dev(np.float64([0.4])), "theta2": dev(np.float32([0.5]))})
    assert m[0].dtype == torch.float64
    # integer Cartesian input is promoted to the default float type
    i32 = c.from_cartesian(torch.tensor([[1], [2], [3], [4]], device=DEV))
    assert i32["r"].dtype == torch.get_default_dtype()
    # outputs never alias inputs
    x = dev(rng.uniform(-1, 1, (4, 16)))
    s = c.from_cartesian(x)
    assert all(v.data_ptr() < x.data_ptr() or v.data_ptr() >= x.data_ptr() + x.numel() * 8 for v in s.values())


def test_odd_sizes_and_misaligned_rows():
    """Row lengths that are not multiples of 16 bytes and batch sizes around the tile size."""
    c, t = us.create_standard(9), vo.create_standard(9)
    rng = np.random.default_rng(21)
    for n in (1, 3, 31, 255, 1023, 1025, 4097, 65537, 300001):
        for ndt in (np.float32, np.float64):
            x = rng.uniform(-1, 1, (c.c_ndim, n)).astype(ndt)
            want = vo.from_cartesian(t, x)
            got = c.from_cartesian(dev(x))
            if n >= 2:
                assert np.array_equal(host(got["r"]), want["r"]), (n, ndt)
            for k in c.s_nodes:
                assert_angles_close(host(got[k]), want[k], f"n={n} {k}")
            back = host(c.to_cartesian({k: dev(v) for k, v in want.items()}, as_array=True))
            assert_leaves_close(back, vo.to_cartesian(t, want, as_array=True), want["r"], depth_of(c), f"n={n}")


def test_stream_and_device_handling():
    c = us.create_hopf(2)
    x = torch.rand((4, 100000), device=DEV, dtype=torch.float64)
    ref = c.from_cartesian(x)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        got = c.from_cartesian(x)
        back = c.to_cartesian(got, as_array=True)
    s.synchronize()
    for k in ref:
        assert torch.equal(ref[k], got[k])
    assert torch.allclose(back, x, atol=1e-14)


def test_more_than_2_pow_31_points():
    """Element offsets past 2^31 (a 17 GB array per leaf set): 64-bit indexing in the tile kernels,
    the ragged tail and the strided kernels.  Checked against the oracle on windows around the
    start, the 2^31 boundary and the end, and by the round trip over the whole batch."""
    free, _ = torch.cuda.mem_get_info()
    if free < 80 << 30:
        pytest.skip("needs 80 GB of free device memory")
    c, t = us.create_polar(), vo.make("polar")
    n = (1 << 31) + 4100  # rows stay 16-byte aligned (tile kernels) and leave a ragged tail of 4 points
    x = torch.empty((2, n), device=DEV, dtype=torch.float32)
    gen = torch.Generator(device=DEV).manual_seed(77)
    for lo in range(0, n, 1 << 28):  # fill in pieces: rand() of 2^32 elements at once is not needed
        hi = min(n, lo + (1 << 28))
        x[:, lo:hi] = torch.rand((2, hi - lo), device=DEV, dtype=torch.float32, generator=gen) * 2 - 1
    s = c.from_cartesian(x)
    assert s["r"].shape == (n,)
    windows = [(0, 5000), ((1 << 31) - 3000, (1 << 31) + 3000), (n - 5000, n)]
    for lo, hi in windows:
        want = vo.from_cartesian(t, host(x[:, lo:hi]))
        assert np.array_equal(host(s["r"][lo:hi]), want["r"]), (lo, hi)
        for k in c.s_nodes:
            assert_angles_close(host(s[k][lo:hi]), want[k], f"window {lo}")
    back = c.to_cartesian(s, as_array=True)
    worst = 0.0
    for lo in range(0, n, 1 << 28):
        hi = min(n, lo + (1 << 28))
        worst = max(worst, (back[:, lo:hi] - x[:, lo:hi]).abs().max().item())
    assert worst <= 1e-6, worst  # a few float32 ulp of |x| <= 1.41 over 2.1e9 points
    # the strided kernels on the same size: a stride-2 view of a wider buffer is not tile-eligible
    del back, s
    torch.cuda.empty_cache()
    y = x[:, 1:]  # misaligned rows (offset by one element): strided path over 2^31 + 4099 points
    s2 = c.from_cartesian(y)
    for lo, hi in [((1 << 31) - 2000, (1 << 31) + 2000), (n - 3001, n - 1)]:
        want = vo.from_cartesian(t, host(y[:, lo:hi]))
        assert np.array_equal(host(s2["r"][lo:hi]), want["r"]), (lo, hi)
        for k in c.s_nodes:
            assert_angles_close(host(s2[k][lo:hi]), want[k], f"strided window {lo}")


def test_concurrent_host_threads_and_streams():
    """The library keeps no per-call global state (parameter blocks travel by value, the few
    static blocks are mutex-guarded): host threads on their own streams, different trees, grids
    and integrals running at once give the bits of a serial run."""
    import threading

    specs = ["standard:9", "hopf:3", "random:32:0", "cbaa", "standard:4", "spherical"]
    trees = [make(us, s) if ":" in s or s == "spherical" else us.create_from_branching_types(s) for s in specs]
    gen = torch.Generator(device=DEV).manual_seed(5)
    xs = [torch.rand((c.c_ndim, 200_003 + 1024 * i), device=DEV, dtype=torch.float64 if i % 2 else torch.float32, generator=gen) * 2 - 1 for i, c in enumerate(trees)]

    def work(c, x):
        s = c.from_cartesian(x)
        back = c.to_cartesian(s, as_array=True)
        jac = c.jacobian(s)
        if c.s_ndim > 9:  # a tensor grid over 32 angles does not exist
            return [*s.values(), back, jac]
        ang, wts = us.roots(c, 5, expand_dims_x=True, xp=torch, device=DEV, dtype=torch.float64)
        grid = c.to_cartesian(ang, as_array=True)
        area = us.integrate(c, lambda a: 1 + 0 * c.to_cartesian(a, as_array=True)[0], False, 5, xp=torch, device=DEV, dtype=torch.float64)
        return [*s.values(), back, jac, grid, area]

    serial = [work(c, x) for c, x in zip(trees, xs)]
    torch.cuda.synchronize()
    results, errors = [None] * len(trees), []

    def run(i):
        try:
            st = torch.cuda.Stream()
            st.wait_stream(torch.cuda.default_stream())
            with torch.cuda.stream(st):
                for _ in range(6):
                    out = work(trees[i], xs[i])
            st.synchronize()
            results[i] = out
        except Exception as e:  # noqa: BLE001
            errors.append((i, repr(e)))

    threads = [threading.Thread(target=run, args=(i,)) for i in range(len(trees))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    for i, (a, b) in enumerate(zip(serial, results)):
        for u, v in zip(a, b):
            assert torch.equal(u.nan_to_num(), v.nan_to_num()), specs[i]


@pytest.mark.parametrize("spec,ndt,n", [("standard:9", np.float32, 1 << 22), ("hopf:3", np.float64, 1 << 21), ("random:32:0", np.float32, 1 << 20)])
def test_million_point_parity(spec, ndt, n):
    """The first 2^20+ points of the benchmark configs against the oracle."""
    c, t = make(us, spec), vo.make(spec)
    x = np.random.default_rng(0).uniform(-1, 1, (c.c_ndim, n)).astype(ndt)
    want = vo.from_cartesian(t, x)
    got = c.from_cartesian(dev(x))
    assert np.array_equal(host(got["r"]), want["r"])
    for k in c.s_nodes:
        assert_angles_close(host(got[k]), want[k], f"{spec} {k}")
    back = host(c.to_cartesian({k: dev(v) for k, v in want.items()}, as_array=True))
    assert_leaves_close(back, vo.to_cartesian(t, want, as_array=True), want["r"], depth_of(c), spec)


@pytest.mark.parametrize(
    "spec,tdt,n",
    [("standard:9", torch.float32, 256_000_000), ("hopf:3", torch.float64, 128_000_000),
     ("standard_prime:8", torch.float64, 128_000_000), ("random:32:0", torch.float32, 64_000_000)],
)
def test_full_size_properties(spec, tdt, n):
    """BASELINE.json sizes, checked through size-independent properties: r equals the norm that
    torch computes in float64, the round trip returns the input, angles lie in their type's
    range, and a re-run is bit-identical (deterministic kernels)."""
    c = make(us, spec)
    g = torch.Generator(device=DEV).manual_seed(1234)
    x = torch.rand((c.c_ndim, n), device=DEV, dtype=tdt, generator=g).mul_(2).sub_(1)
    sph = c.from_cartesian(x)
    eps = torch.finfo(tdt).eps
    chunk = 1 << 24
    for lo in range(0, n, chunk):
        sl = slice(lo, min(n, lo + chunk))
        r64 = x[:, sl].double().square().sum(0).sqrt()
        assert ((sph["r"][sl].double() - r64).abs() <= c.c_ndim * eps * r64).all()
    pi = float(np.pi)
    for node in c.s_nodes:
        lo_, hi_ = {"a": (-pi, pi), "b": (0, pi), "b'": (-pi / 2, pi / 2), "c": (0, pi / 2)}[c.branching_types[node].value]
        th = sph[node]
        assert th.min().item() >= lo_ - 4 * eps and th.max().item() <= hi_ + 4 * eps, node
    back = c.to_cartesian(sph, as_array=True)
    for lo in range(0, n, chunk):
        sl = slice(lo, min(n, lo + chunk))
        err = (back[:, sl] - x[:, sl]).abs().amax(0) / sph["r"][sl]
        assert err.max().item() <= 8 * eps
    sph2 = c.from_cartesian(x)
    assert all(torch.equal(sph[k], sph2[k]) for k in sph)
    del back, sph2
    torch.cuda.empty_cache()


def test_custom_ops_match_the_methods():
    """torch.ops.ultrasphere_b200.* launch the same kernels as the SphericalCoordinates methods."""
    c = us.create_hopf(3)
    h = us.plan_handle(c)
    x = torch.rand((c.c_ndim, 50_000), device=DEV, dtype=torch.float64) * 2 - 1
    s = c.from_cartesian(x)
    packed = torch.ops.ultrasphere_b200.from_cartesian([x[i] for i in range(c.c_ndim)], h)
    assert torch.equal(packed[0], s["r"])
    for i, node in enumerate(c.s_nodes):
        assert torch.equal(packed[1 + i], s[node])
    back = torch.ops.ultrasphere_b200.to_cartesian([s[node] for node in c.s_nodes], s["r"], h)
    assert torch.equal(back, c.to_cartesian(s, as_array=True))
    no_r = torch.ops.ultrasphere_b200.to_cartesian([s[node] for node in c.s_nodes], None, h)
    assert torch.equal(no_r, c.to_cartesian({k: v for k, v in s.items() if k != "r"}, as_array=True))
    val = torch.rand((6, 7, 3), device=DEV, dtype=torch.float64)
    w = [torch.rand(6, device=DEV, dtype=torch.float64), torch.rand(7, device=DEV, dtype=torch.float64)]
    assert torch.equal(torch.ops.ultrasphere_b200.weighted_reduce(val, w), us.weighted_reduce(val, w))
    torch.library.opcheck(torch.ops.ultrasphere_b200.from_cartesian.default, ([x[i] for i in range(c.c_ndim)], h), test_utils=("test_schema", "test_faketensor"))


@pytest.mark.parametrize("spec,tdt", [("standard:9", torch.float32), ("hopf:3", torch.float32), ("hopf:3", torch.float64), ("random:6:3", torch.float64)])
def test_tiled_and_strided_paths_give_identical_bits(spec, tdt):
    """Large batches run the TMA-tiled kernels, small ones the strided kernels; both must produce
    the same bits (same device math, no FMA contraction in the norm chains of either)."""
    c = make(us, spec)
    n = 1 << 20
    x = torch.rand((c.c_ndim, n), device=DEV, dtype=tdt) * 2 - 1
    whole = c.from_cartesian(x)
    back_whole = c.to_cartesian(whole, as_array=True)
    step = 4099  # far below the tile threshold and not a multiple of the vector width
    for lo in range(0, n, step * 37):
        hi = min(n, lo + step)
        part = c.from_cartesian(x[:, lo:hi].contiguous())
        for k in whole:
            assert torch.equal(part[k], whole[k][lo:hi]), (spec, k, lo)
        back = c.to_cartesian({k: v[lo:hi].contiguous() for k, v in whole.items()}, as_array=True)
        assert torch.equal(back, back_whole[:, lo:hi])


@pytest.mark.parametrize("dname,ndt,tdt", DTYPES)
@pytest.mark.parametrize("spec", ["standard:9", "standard_prime:8", "hopf:3", "random:32:0"])
def test_tiled_ring_and_ragged_tail_against_oracle(spec, dname, ndt, tdt):
    """Batch sizes that run the tile / chain / ring kernels plus a ragged tail (< 1 tile) handled by
    the strided kernel, and an odd size whose rows are not 16-byte aligned (strided throughout)."""
    c, t = make(us, spec), vo.make(spec)
    rng = np.random.default_rng(17)
    for n in ((1 << 20) + 772, (1 << 19) + 1):
        x = rng.uniform(-1, 1, (c.c_ndim, n)).astype(ndt)
        want = vo.from_cartesian(t, x)
        got = c.from_cartesian(dev(x))
        assert np.array_equal(host(got["r"]), want["r"]), (spec, n)
        for k in c.s_nodes:
            assert_angles_close(host(got[k]), want[k], f"{spec} n={n} {k}")
        back = host(c.to_cartesian({k: dev(v) for k, v in want.items()}, as_array=True))
        assert_leaves_close(back, vo.to_cartesian(t, want, as_array=True), want["r"], depth_of(c), f"{spec} n={n}")
        no_r = host(c.to_cartesian({k: dev(want[k]) for k in c.s_nodes}, as_array=True))
        wanted = vo.to_cartesian(t, {k: want[k] for k in c.s_nodes}, as_array=True)
        assert_leaves_close(no_r, wanted, np.ones(1, ndt), depth_of(c), f"{spec} n={n} no r")


def test_cuda_graph_capture_and_replay():
    """The launches are plain stream-ordered kernels with no host synchronisation, so a round trip
    can be captured once in a CUDA graph and replayed on new data (launch-bound small batches)."""
    c = us.create_standard(9)
    n = 4096
    static_x = torch.rand((c.c_ndim, n), device=DEV) * 2 - 1
    c.to_cartesian(c.from_cartesian(static_x), as_array=True)  # warm-up: plan, caches, allocator
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        static_s = c.from_cartesian(static_x)
        static_y = c.to_cartesian(static_s, as_array=True)
    for seed in (1, 2, 3):
        fresh = torch.rand((c.c_ndim, n), device=DEV, generator=torch.Generator(device=DEV).manual_seed(seed)) * 2 - 1
        static_x.copy_(fresh)
        graph.replay()
        torch.cuda.synchronize()
        eager = c.from_cartesian(fresh)
        for k in eager:
            assert torch.equal(static_s[k], eager[k])
        assert torch.equal(static_y, c.to_cartesian(eager, as_array=True))


@pytest.mark.parametrize("spec,ndt", [("standard:256", np.float64), ("hopf:8", np.float64), ("hopf:8", np.float32), ("random:200:5", np.float64)])
def test_maximum_tree_sizes(spec, ndt):
    """Capacity limits of the compiled plan: 256 angles (a 257-D chain, leaf depth 256), the perfect
    binary tree with 255 angles (deferred-branch stack depth 7) and a 200-angle random tree."""
    c, t = make(us, spec), vo.make(spec)
    assert c.s_ndim <= 256
    n = 2048
    x = np.random.default_rng(23).uniform(-1, 1, (c.c_ndim, n)).astype(ndt)
    want = vo.from_cartesian(t, x)
    got = c.from_cartesian(dev(x))
    assert list(got) == list(want)
    assert np.array_equal(host(got["r"]), want["r"])
    for k in c.s_nodes:
        assert_angles_close(host(got[k]), want[k], f"{spec} {k}")
    back = host(c.to_cartesian({k: dev(v) for k, v in want.items()}, as_array=True))
    assert_leaves_close(back, vo.to_cartesian(t, want, as_array=True), want["r"], depth_of(c), spec)
    with pytest.raises(ValueError):
        us.create_standard(257).from_cartesian(torch.zeros((258, 4), device=DEV))


@pytest.mark.parametrize("spec", ["spherical", "standard:4", "hopf:3", "expr:cbabba", "random:32:0"])
def test_jacobian_is_the_intended_formula(spec):
    """the reference's jacobian() is NaN for every tree (leaf entries of S are NaN);
    ours implements the formula it intends.  Checked against a float64 NumPy evaluation and, for the
    standard tree, against the textbook sin-power form and the sphere's area."""
    c, t = make(us, spec), vo.make(spec)
    rng = np.random.default_rng(31)
    n = 5000
    ang = {k: rng.uniform(0.1, 1.4, n) for k in c.s_nodes}
    rr = rng.uniform(0.5, 2.0, n)
    want = rr.copy()
    for node in t.s_nodes:
        for kind, fn in (("cos", np.cos), ("sin", np.sin)):
            child = t.child(node, kind)
            power = 0 if not t.adj[child] else int(t.S[child]) + 1
            want = want * fn(ang[node]) ** power
    got = c.jacobian({k: dev(v) for k, v in ang.items()} | {"r": dev(rr)})
    np.testing.assert_allclose(host(got), want, rtol=1e-13)
    no_r = c.jacobian({k: dev(v) for k, v in ang.items()})
    np.testing.assert_allclose(host(no_r), want / rr, rtol=1e-13)
    f32 = c.jacobian({k: dev(v.astype(np.float32)) for k, v in ang.items()})
    assert f32.dtype == torch.float32
    # float32: every factor c^p carries ~p ulp; deep random trees have p up to 32 and may underflow
    np.testing.assert_allclose(host(f32), want / rr, rtol=5e-4, atol=1e-30)


def test_jacobian_standard_tree_and_broadcast():
    c = us.create_standard(3)  # J = sin^2(theta0) * sin(theta1)
    th = {k: torch.rand(1000, device=DEV, dtype=torch.float64) * 3 for k in c.s_nodes}
    want = torch.sin(th["theta0"]) ** 2 * torch.sin(th["theta1"])
    assert torch.allclose(c.jacobian(th), want, rtol=1e-13, atol=0)
    # on the quadrature grid (broadcast 1-D axes): integrating J over the angle box gives the area
    n = 40
    ax = {"theta0": torch.linspace(0, torch.pi, n + 1, device=DEV, dtype=torch.float64)[:, None, None],
          "theta1": torch.linspace(0, torch.pi, n + 1, device=DEV, dtype=torch.float64)[None, :, None],
          "theta2": torch.linspace(0, 2 * torch.pi, 2 * n + 1, device=DEV, dtype=torch.float64)[None, None, :]}
    J = c.jacobian(ax)
    assert J.shape == (n + 1, n + 1, 2 * n + 1)
    area = torch.trapezoid(torch.trapezoid(torch.trapezoid(J, ax["theta2"].ravel(), dim=2), ax["theta1"].ravel(), dim=1), ax["theta0"].ravel(), dim=0)
    assert abs(area.item() - c.surface_area()) < 2e-2


def test_row_matrix_entry_points_and_plan_seal():
    """us_from_cartesian_rows / us_to_cartesian_rows (base pointer + row pitch) give the bits of the
    pointer-table entry points, and every entry point refuses a plan whose seal does not match
    (a hand-built or modified us_plan_t must not reach the kernels)."""
    import ctypes as C

    from ultrasphere_b200 import _native
    from ultrasphere_b200._plan import plan_for

    lib = _native.lib()
    c = us.create_hopf(3)
    ct = plan_for(c)
    n = 300_000
    x = torch.rand((c.c_ndim, n), device=DEV, dtype=torch.float64) * 2 - 1
    want = c.from_cartesian({k: x[i].clone() for i, k in enumerate(c.c_nodes)})  # mapping input: pointer tables
    out = torch.empty((c.s_ndim + 1, n), device=DEV, dtype=torch.float64)
    stream = torch.cuda.current_stream().cuda_stream
    assert lib.us_from_cartesian_rows(ct.plan_ref, _native.US_F64, n, x.data_ptr(), n * 8, out.data_ptr(), n * 8, stream) == 0
    assert torch.equal(out[0], want["r"])
    for s, node in enumerate(c.s_nodes):
        assert torch.equal(out[1 + s], want[node]), node
    back = torch.empty_like(x)
    ang = (C.c_void_p * c.s_ndim)(*[out[1 + s].data_ptr() for s in range(c.s_ndim)])
    assert lib.us_to_cartesian_rows(ct.plan_ref, _native.US_F64, n, ang, out[0].data_ptr(), back.data_ptr(), n * 8, stream) == 0
    assert torch.equal(back, c.to_cartesian(want, as_array=True))
    forged = _native.UsPlan.from_buffer_copy(ct.plan)
    forged.node[0] ^= 1 << 24  # another leaf slot: still plausible counts, different program
    rc = lib.us_from_cartesian_rows(C.byref(forged), _native.US_F64, n, x.data_ptr(), n * 8, out.data_ptr(), n * 8, stream)
    assert rc == -3 and b"us_plan_compile" in lib.us_last_error_string()
    torch.cuda.synchronize()


@pytest.mark.parametrize("spec,tdt", [("standard:9", torch.float32), ("hopf:3", torch.float64), ("spherical", torch.float64)])
def test_pool_geometry_does_not_change_the_bits(spec, tdt):
    """us_set_tuning only moves work between consumer groups and buffers: whatever the cut of the
    shared-memory pool (and with the ragged last tile riding in the same launch), the results are
    bit-identical."""
    from ultrasphere_b200 import _native

    lib = _native.lib()
    c = make(us, spec)
    n = (1 << 21) + 1284  # ragged: the last tile is partial, a multiple of 16 bytes
    x = torch.rand((c.c_ndim, n), device=DEV, dtype=tdt) * 2 - 1
    ref_s = c.from_cartesian(x)
    ref_x = c.to_cartesian(ref_s, as_array=True)
    try:
        for groups, bufs, u in ((1, 2, 0), (2, 3, 1), (3, 7, 2), (5, 6, 1), (6, 11, 0), (4, 5, 2)):
            assert lib.us_set_tuning(b"tile_groups", groups) == 0 and lib.us_set_tuning(b"tile_bufs", bufs) == 0
            assert lib.us_set_tuning(b"tile_u", u) == 0
            s = c.from_cartesian(x)
            for k in ref_s:
                assert torch.equal(s[k], ref_s[k]), (groups, bufs, k)
            assert torch.equal(c.to_cartesian(s, as_array=True), ref_x), (groups, bufs)
        assert lib.us_set_tuning(b"no_such_knob", 1) == -1
    finally:
        lib.us_set_tuning(b"tile_groups", 0)
        lib.us_set_tuning(b"tile_bufs", 0)
        lib.us_set_tuning(b"tile_u", 0)


@pytest.mark.parametrize("spec,tdt", [("random:32:0", torch.float32), ("random:32:0", torch.float64), ("hopf:5", torch.float32)])
def test_ring_geometry_does_not_change_the_bits(spec, tdt):
    """The row-ring kernels of the wide trees: one or two 16-byte chunks of a row per thread, two to
    four CTAs per SM (hence ring depths from 4 to 13 slots) — bit-identical results, also against
    the strided kernels (tile_max_rows = 0 sends every tree through the ring; a small batch through
    the strided path).  create_hopf(5) (31 angles, stack depth 4) parks the most rows."""
    from ultrasphere_b200 import _native

    lib = _native.lib()
    c = make(us, spec)
    n = 296 * 2048 * 2 + 4 * 321  # enough ring tiles for every geometry, ragged tail for the strided kernel
    x = torch.rand((c.c_ndim, n), device=DEV, dtype=tdt) * 2 - 1
    m = 40_000  # below the ring threshold: strided kernels
    small_s = c.from_cartesian(x[:, :m].contiguous())
    try:
        ref_s = ref_x = None
        for u, ctas, slots in ((0, 0, 0), (1, 0, 0), (1, 2, 0), (2, 2, 5), (2, 3, 0), (1, 4, 4), (2, 2, 32)):
            assert lib.us_set_tuning(b"ring_u", u) == 0 and lib.us_set_tuning(b"ring_ctas", ctas) == 0
            assert lib.us_set_tuning(b"ring_slots", slots) == 0
            s = c.from_cartesian(x)
            back = c.to_cartesian(s, as_array=True)
            if ref_s is None:
                ref_s, ref_x = s, back
                for k in s:
                    assert torch.equal(s[k][:m], small_s[k]), ("strided", k)
                continue
            for k in ref_s:
                assert torch.equal(s[k], ref_s[k]), (u, ctas, slots, k)
            assert torch.equal(back, ref_x), (u, ctas, slots)
    finally:
        lib.us_set_tuning(b"ring_u", 0)
        lib.us_set_tuning(b"ring_ctas", 0)
        lib.us_set_tuning(b"ring_slots", 0)


@pytest.mark.parametrize("spec,tdt,n,knobs", [
    ("random:32:0", torch.float32, 64_000_000, {}),
    ("random:32:0", torch.float32, 64_000_000, {b"ring_slots": 4, b"ring_ctas": 3}),
    ("random:32:0", torch.float64, 32_000_000, {}),
    ("standard:9", torch.float32, 256_000_000, {}),
    ("hopf:3", torch.float64, 128_000_000, {}),
    ("standard_prime:8", torch.float64, 128_000_000, {b"tile_u": 2}),
])
def test_back_to_back_launches_are_deterministic(spec, tdt, n, knobs):
    """from_cartesian and to_cartesian launched back to back, over and over, at the BASELINE sizes:
    every repetition must reproduce the first one bit for bit.  This is the test that catches a
    shared-memory slot handed back to the TMA producer too early (a refill overtaking a queued load
    shows up as a few dozen wrong points in ~10^9, only with shallow rings and only now and then)."""
    from ultrasphere_b200 import _native

    lib = _native.lib()
    c = make(us, spec)
    x = torch.rand((c.c_ndim, n), device=DEV, dtype=tdt, generator=torch.Generator(device=DEV).manual_seed(99)).mul_(2).sub_(1)
    try:
        for k, v in knobs.items():
            assert lib.us_set_tuning(k, v) == 0
        s0 = c.from_cartesian(x)
        y0 = c.to_cartesian(s0, as_array=True)
        for rep in range(25):
            s = c.from_cartesian(x)
            y = c.to_cartesian(s, as_array=True)
            assert torch.equal(y, y0), (spec, rep, int((y != y0).any(0).sum()))
            for k in s0:
                assert torch.equal(s[k], s0[k]), (spec, rep, k)
    finally:
        for k in knobs:
            lib.us_set_tuning(k, 0)
        torch.cuda.empty_cache()


def test_parity_report():
    """The measured error distribution behind the parity claim, per BASELINE config and dtype, on
    2^20 seeded points each: max and 99.99th-percentile distance from the NumPy oracle for r, the
    angles (ulp of the angle) and the leaves (ulp element-wise and in ulp(r)), float32 also as a
    relative error.  Written to gpurun_out/parity_report.json (copied to profiles/ by the builder);
    the assertions are the north star's bars."""
    import json
    import os

    n = 1 << 20
    report = {"points_per_config": n, "oracle": "oracle/vks_oracle.py (NumPy), inputs default_rng(0).uniform(-1, 1)", "configs": {}}
    for spec in ("spherical", "standard:9", "hopf:3", "standard_prime:8", "random:32:0"):
        c, t = make(us, spec), vo.make(spec)
        depth = depth_of(c)
        for dname, ndt, tdt in DTYPES:
            x = np.random.default_rng(0).uniform(-1, 1, (c.c_ndim, n)).astype(ndt)
            want = vo.from_cartesian(t, x)
            got = c.from_cartesian(dev(x))
            eps = np.finfo(ndt).eps
            d_r = vo.ulp_distance(host(got["r"]), want["r"])
            ang_g = np.stack([host(got[k]) for k in c.s_nodes])
            ang_w = np.stack([want[k] for k in c.s_nodes])
            d_a = vo.ulp_distance(ang_g, ang_w)
            back = host(c.to_cartesian({k: dev(v) for k, v in want.items()}, as_array=True))
            wantx = vo.to_cartesian(t, want, as_array=True)
            d_x = vo.ulp_distance(back, wantx)
            x_over_r = np.abs(back.astype(np.float64) - wantx.astype(np.float64)) / np.spacing(np.abs(want["r"])).astype(np.float64)[None]  # ulp of r in the data's own type
            rt = host(c.to_cartesian(got, as_array=True))
            entry = {
                "r_max_ulp": float(d_r.max()),
                "angle_max_ulp": float(d_a.max()), "angle_p9999_ulp": float(np.quantile(d_a, 0.9999)),
                "leaf_max_ulp": float(d_x.max()), "leaf_p9999_ulp": float(np.quantile(d_x, 0.9999)), "leaf_frac_over_4ulp": float((d_x > 4).mean()),
                "leaf_max_err_in_ulp_of_r": float(x_over_r.max()),
                "max_leaf_depth": int(depth.max()),
                "round_trip_max_abs_over_r_eps": float(np.max(np.abs(rt.astype(np.float64) - x) / want["r"].astype(np.float64)[None]) / eps),
            }
            if ndt == np.float32:
                with np.errstate(all="ignore"):
                    entry["angle_max_rel"] = float(np.nanmax(np.abs(ang_g.astype(np.float64) - ang_w) / np.abs(ang_w.astype(np.float64))))
                    rel = np.abs(back.astype(np.float64) - wantx) / np.abs(wantx.astype(np.float64))
                    entry["leaf_max_rel"] = float(np.nanmax(np.where(np.isfinite(rel), rel, 0)))
            report["configs"][f"{spec}|{dname}"] = entry
            assert entry["r_max_ulp"] == 0, (spec, dname)
            if ndt == np.float64:
                assert entry["angle_max_ulp"] <= F64_ULP and entry["leaf_max_err_in_ulp_of_r"] <= F64_ULP, (spec, entry)
                assert entry["leaf_max_ulp"] <= F64_ULP + depth.max(), (spec, entry)
            else:
                assert entry["angle_max_rel"] <= F32_REL, (spec, entry)
    out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    with open(os.path.join(out_dir, "parity_report.json"), "w") as fh:
        json.dump(report, fh, indent=1)
