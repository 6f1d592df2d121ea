"""us_to_cartesian_dot: sum_v k[v] * x_v fused into the transform, against the oracle's
``einsum("v,v...->...", k, to_cartesian(s, as_array=True))`` (the reference's own quadrature
integrand, tests/test_integral.py:88-92) on dense batches, tensor grids and inside integrate()."""
import os

import numpy as np
import pytest
import torch

import ultrasphere_b200 as us
from oracle import vks_oracle as vo
from tests.helpers import GOLDEN, INTEGRAL_REL, make
from tests.test_gpu_parity import depth_of, dev, host

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
QD = np.load(os.path.join(GOLDEN, "quadrature.npz"))
SPECS = ["polar", "spherical", "standard:4", "standard:9", "standard_prime:3", "hopf:2", "hopf:3", "random:6:3", "random:32:0", "expr:cbabba"]
DTYPES = [("f32", np.float32, torch.float32), ("f64", np.float64, torch.float64)]


def assert_dot_close(got, k, x, depth, what=""):
    """got vs sum_v k[v] x[v] (x: oracle leaves, float64 accumulate); every leaf carries
    (4 + depth) ulp of its own, the FMA chain one more per term."""
    want = np.einsum("v,v...->...", k.astype(np.float64), x.astype(np.float64))
    eps = np.finfo(got.dtype).eps
    weight = (8.0 + depth).reshape((-1,) + (1,) * (x.ndim - 1))
    bound = eps * np.sum(weight * np.abs(k.astype(np.float64)).reshape(weight.shape) * np.abs(x.astype(np.float64)), axis=0)
    err = np.abs(got.astype(np.float64) - want)
    assert got.shape == want.shape, (what, got.shape, want.shape)
    assert (err <= bound + np.finfo(got.dtype).tiny).all(), f"{what}: worst {np.max(err / (bound + 1e-300))} of the bound"


@pytest.mark.parametrize("dname,ndt,tdt", DTYPES)
@pytest.mark.parametrize("spec", SPECS)
def test_dense_batches_match_oracle(spec, dname, ndt, tdt):
    c, t = make(us, spec), vo.make(spec)
    rng = np.random.default_rng(5)
    d = depth_of(c)
    for n in (1, 7, 4096, 50001):
        ang = {node: rng.uniform(-3.2, 3.2, n).astype(ndt) for node in c.s_nodes}
        k = rng.normal(size=c.c_ndim).astype(ndt)
        for with_r in (False, True):
            a = dict(ang)
            if with_r:
                a["r"] = rng.uniform(0.2, 3.0, n).astype(ndt)
            x = vo.to_cartesian(t, dict(a), as_array=True)
            got = c.to_cartesian_dot({key: dev(v) for key, v in a.items()}, dev(k))
            assert got.dtype == tdt
            assert_dot_close(host(got), k, x, d, f"{spec} n={n} r={with_r}")


@pytest.mark.parametrize("spec", sorted({key.split("|")[0] for key in QD.files}))
def test_tensor_grids_match_oracle_and_the_unfused_form(spec):
    c, t = make(us, spec), vo.make(spec)
    rng = np.random.default_rng(11)
    d = depth_of(c)
    for n in (3, 8):  # odd innermost extent: scalar stores; even: 16-byte stores
        xs, _ = us.roots(c, n, expand_dims_x=True, xp=torch, device=DEV, dtype=torch.float64)
        k = rng.normal(size=(3, c.c_ndim))
        grid = vo.to_cartesian(t, {key: host(v) for key, v in xs.items()}, as_array=True)
        got = c.to_cartesian_dot(xs, torch.from_numpy(k))  # k on the host, matrix form
        assert got.shape == (3, *grid.shape[1:])
        unfused = torch.einsum("mv,v...->m...", torch.from_numpy(k).to(DEV), c.to_cartesian(xs, as_array=True))
        for j in range(3):
            assert_dot_close(host(got[j]), k[j], grid, d, f"{spec} n={n} row {j}")
        assert (got - unfused).abs().max().item() <= 1e-13 * max(1.0, float(np.abs(k).sum()))
        # with a radius axis in front and a python-list k
        s = {key: v[None] for key, v in xs.items()}
        s["r"] = torch.linspace(0.5, 2.0, 4, device=DEV, dtype=torch.float64).reshape((4,) + (1,) * c.s_ndim)
        got_r = c.to_cartesian_dot(s, list(k[0]))
        grid_r = vo.to_cartesian(t, {key: host(v) for key, v in s.items()}, as_array=True)
        assert_dot_close(host(got_r), k[0], grid_r, d, f"{spec} n={n} with r")


def test_shapes_arguments_and_errors():
    c = us.create_hopf(2)
    th = {node: torch.rand(5, device=DEV, dtype=torch.float64) for node in c.s_nodes}
    with pytest.raises(ValueError):
        c.to_cartesian_dot(th, [1.0, 2.0])
    with pytest.raises(KeyError):
        c.to_cartesian_dot({"theta0": th["theta0"]}, [1.0] * 4)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        c.to_cartesian_dot({key: v.cpu() for key, v in th.items()}, [1.0] * 4)
    one = c.to_cartesian_dot(th, [1.0, 0.0, 0.0, 0.0])
    assert torch.allclose(one, c.to_cartesian(th)[c.c_nodes[0]], rtol=0, atol=1e-15)
    mixed = c.to_cartesian_dot({**{key: v.float() for key, v in th.items()}, "r": 2.0}, torch.ones(4))
    assert mixed.dtype == torch.float32 and mixed.shape == (5,)
    empty = c.to_cartesian_dot({node: torch.empty((0, 3), device=DEV) for node in c.s_nodes}, torch.ones((2, 4)))
    assert empty.shape == (2, 0, 3)
    z = us.create_standard(0)
    assert torch.equal(z.to_cartesian_dot({"r": torch.tensor([1.0, -2.0], device=DEV)}, [3.0]), torch.tensor([3.0, -6.0], device=DEV))


@pytest.mark.parametrize("spec", sorted({key.split("|")[0] for key in QD.files}))
def test_integrals_with_the_fused_integrand_match_reference_golden(spec):
    c = make(us, spec)
    k = torch.from_numpy(QD[f"{spec}|k"]).to(DEV)
    for n in (4, 7, 12):
        if f"{spec}|n{n}|dense" not in QD.files:
            continue
        want = float(QD[f"{spec}|n{n}|dense"])
        got = us.integrate(c, lambda s: torch.exp(c.to_cartesian_dot(s, k)), False, n, xp=torch, device=DEV, dtype=torch.float64)
        assert abs(got.item() - want) <= INTEGRAL_REL * abs(want), (spec, n)


def test_full_size_quadrature_with_the_fused_integrand():
    """C5 (create_standard(4), n = 96, 170 M nodes): exp(k.x) through to_cartesian_dot, slabbed and whole."""
    k5 = np.random.default_rng(0).uniform(0, 1, 5)
    kd = torch.from_numpy(k5).to(DEV)
    c = us.create_standard(4)
    want = vo.analytic_exp_kx(k5)
    for slab in (None, 64 << 20):
        got = us.integrate(c, lambda s: torch.exp(c.to_cartesian_dot(s, kd)), False, 96, xp=torch, device=DEV, dtype=torch.float64, slab_bytes=slab).item()
        assert abs(got - want) <= INTEGRAL_REL * want


# ---- the row kernel (chain trees on grids with long rows) ------------------------------------------
def _both_kernels(fn):
    """Run fn() through the row kernel and, with US_NO_ROWS set, through the generic grid kernel."""
    rows = fn()
    os.environ["US_NO_ROWS"] = "1"
    try:
        generic = fn()
    finally:
        del os.environ["US_NO_ROWS"]
    return rows, generic


ROW_CASES = [
    # spec, per-angle extents (s_nodes order), which launch axis each angle lies on
    ("standard:3", (64, 64, 128), (0, 1, 2)),    # integrate()'s layout, 128-point rows
    ("standard:3", (50, 60, 200), (0, 1, 2)),    # row length not a multiple of the warp chunk
    ("standard:3", (24, 100, 520), (0, 1, 2)),   # rows longer than one chunk of 32 * U * V points
    ("standard:3", (96, 40, 64), (2, 0, 1)),     # the ROOT angle varies along the row: vector mode from the start
    ("standard_prime:3", (64, 48, 96), (0, 1, 2)),
    ("spherical", (2500, 136), (0, 1)),
    ("standard:4", (24, 24, 24, 48), (0, 1, 2, 3)),
]


@pytest.mark.parametrize("dname,ndt,tdt", DTYPES)
@pytest.mark.parametrize("spec,extents,axes", ROW_CASES)
def test_row_kernel_matches_oracle_and_generic_kernel(spec, extents, axes, dname, ndt, tdt):
    c, t = make(us, spec), vo.make(spec)
    rng = np.random.default_rng(23)
    nd = len(extents)
    full = [0] * nd
    for e, a in zip(extents, axes):
        full[a] = e
    ang = {}
    for node, e, a in zip(c.s_nodes, extents, axes):
        shp = [1] * nd
        shp[a] = e
        ang[node] = rng.uniform(-3.2, 3.2, shp).astype(ndt)
    k = rng.normal(size=c.c_ndim).astype(ndt)
    d = depth_of(c)
    r_variants = [None, 1.5, rng.uniform(0.5, 2.0, [full[0]] + [1] * (nd - 1)).astype(ndt), rng.uniform(0.5, 2.0, [1] * (nd - 1) + [full[-1]]).astype(ndt)]
    for r in r_variants:
        a_np = dict(ang)
        a_dev = {key: dev(v) for key, v in ang.items()}
        if r is not None:
            a_np["r"] = r
            a_dev["r"] = r if isinstance(r, float) else dev(r)
        want = vo.to_cartesian(t, dict(a_np), as_array=True).astype(ndt)
        rows, generic = _both_kernels(lambda: c.to_cartesian(dict(a_dev), as_array=True))
        assert torch.equal(rows, generic), (spec, extents, "grid")
        rr = np.broadcast_to(np.asarray(1.0 if r is None else r, dtype=np.float64), want.shape[1:])
        from tests.test_gpu_parity import assert_leaves_close

        assert_leaves_close(host(rows), want, rr, d, f"{spec} {extents}")
        drows, dgeneric = _both_kernels(lambda: c.to_cartesian_dot(dict(a_dev), dev(k)))
        assert torch.equal(drows, dgeneric), (spec, extents, "dot")
        assert_dot_close(host(drows), k, want, d, f"{spec} {extents} dot")


def test_row_kernel_with_a_dense_angle_and_no_table():
    """One angle given on the full grid (no sincos table: evaluated per point inside the row kernel)."""
    c, t = us.create_standard(3), vo.create_standard(3)
    rng = np.random.default_rng(31)
    shape = (40, 70, 96)
    ang = {"theta0": rng.uniform(0, 3, (40, 1, 1)), "theta1": rng.uniform(0, 3, shape), "theta2": rng.uniform(0, 6, (1, 1, 96))}
    a_dev = {key: dev(v) for key, v in ang.items()}
    want = vo.to_cartesian(t, dict(ang), as_array=True)
    rows, generic = _both_kernels(lambda: c.to_cartesian(dict(a_dev), as_array=True))
    assert torch.equal(rows, generic)
    from tests.test_gpu_parity import assert_leaves_close

    assert_leaves_close(host(rows), want, np.ones(1), depth_of(c), "dense angle")
    # the functional through the row kernel: per-point sincos of the dense angle inside the vector phase
    k = rng.normal(size=4)
    drows, dgeneric = _both_kernels(lambda: c.to_cartesian_dot(dict(a_dev), dev(k)))
    assert torch.equal(drows, dgeneric)
    assert_dot_close(host(drows), k, want, depth_of(c), "dense angle dot")
    # ... and with the dense angle at the root, so that every later node is row-uniform in vector mode
    ang2 = {"theta0": rng.uniform(0, 3, shape), "theta1": rng.uniform(0, 3, (1, 70, 1)), "theta2": rng.uniform(0, 6, (40, 1, 1))}
    a2 = {key: dev(v) for key, v in ang2.items()}
    want2 = vo.to_cartesian(t, dict(ang2), as_array=True)
    d2, g2 = _both_kernels(lambda: c.to_cartesian_dot(dict(a2), dev(k)))
    assert torch.equal(d2, g2)
    assert_dot_close(host(d2), k, want2, depth_of(c), "dense root angle dot")
