"""integrate() / roots() on the CUDA path against the oracle and the reference's golden values."""
import json
import math
import os

import numpy as np
import pytest
import torch

import ultrasphere_b200 as us
from oracle import vks_oracle as vo
from tests.helpers import GOLDEN, INTEGRAL_REL, make

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
QD = np.load(os.path.join(GOLDEN, "quadrature.npz"))
QUAD_SPECS = sorted({k.split("|")[0] for k in QD.files})


def host(t):
    return t.detach().cpu().numpy()


@pytest.mark.parametrize("spec", QUAD_SPECS)
def test_roots_are_the_reference_tables(spec):
    c = make(us, spec)
    for n in (4, 7, 12):
        xs, ws = us.roots(c, n, expand_dims_x=False, xp=torch, device=DEV, dtype=torch.float64)
        for i, node in enumerate(c.s_nodes):
            np.testing.assert_allclose(host(xs[node]), QD[f"{spec}|n{n}|x{i}"], rtol=1e-14, atol=1e-15)
            np.testing.assert_allclose(host(ws[node]), QD[f"{spec}|n{n}|w{i}"], rtol=1e-13, atol=1e-300)
        xe, we = us.roots(c, n, expand_dims_x=True, expand_dims_w=True, xp=torch, device=DEV, dtype=torch.float64)
        for i, node in enumerate(c.s_nodes):
            shp = [1] * c.s_ndim
            shp[i] = xs[node].numel()
            assert list(xe[node].shape) == shp and list(we[node].shape) == shp


def test_roots_dtype_quirk_matches_reference():
    """dtype=None: type-a angles come from arange (default dtype), the others from float64 tables."""
    c = us.create_spherical()
    xs, ws = us.roots(c, 5, expand_dims_x=False, xp=torch, device=DEV)
    assert xs["theta"].dtype == torch.float64 and xs["phi"].dtype == torch.get_default_dtype()


@pytest.mark.parametrize("slab", [None, 0, 4096])
@pytest.mark.parametrize("spec", QUAD_SPECS)
def test_integrals_match_reference_golden(spec, slab):
    c = make(us, spec)
    k = torch.from_numpy(QD[f"{spec}|k"]).to(DEV)

    def f(s):
        return torch.exp(torch.einsum("v,v...->...", k, c.to_cartesian(s, as_array=True)))

    def fv(s):
        x = c.to_cartesian(s, as_array=True)
        ph = torch.einsum("v,v...->...", k, x)
        return torch.stack([torch.exp(1j * ph), ph**2 + 0j, torch.exp(ph) + 1j * x[0]], dim=-1)

    def fsep(s):
        return {
            node: torch.stack([torch.cos(s[node]) ** 2, 1.0 + 0 * s[node], torch.sin(s[node])], dim=-1)
            for node in c.s_nodes
        }

    kw = dict(xp=torch, device=DEV, dtype=torch.float64)
    for n in (4, 7, 12):
        if f"{spec}|n{n}|dense" not in QD.files:
            continue
        want = QD[f"{spec}|n{n}|dense"]
        got = us.integrate(c, f, False, n, slab_bytes=slab, **kw)
        assert got.shape == () and got.dtype == torch.float64
        assert abs(got.item() - float(want)) <= INTEGRAL_REL * abs(float(want)), (spec, n)
        wantv = QD[f"{spec}|n{n}|dense_vec"]
        gotv = host(us.integrate(c, fv, False, n, slab_bytes=slab, **kw))
        assert gotv.shape == wantv.shape and gotv.dtype == np.complex128
        scale = np.abs(wantv).max()
        assert np.abs(gotv - wantv).max() <= INTEGRAL_REL * max(scale, 1.0) * 10, (spec, n)
        sep = us.integrate(c, fsep, True, n, **kw)
        assert list(sep) == list(c.s_nodes)
        gots = np.stack([host(sep[node]) for node in c.s_nodes])
        np.testing.assert_allclose(gots, QD[f"{spec}|n{n}|sep"], rtol=INTEGRAL_REL, atol=1e-13)


def test_analytic_and_doctest_known_answers():
    with open(os.path.join(GOLDEN, "kat.json")) as fh:
        kat = json.load(fh)
    c = us.create_spherical()
    val = us.integrate(c, lambda s: s["theta"] ** 2 * s["phi"], False, 10, xp=torch, device=DEV, dtype=torch.float64)
    assert round(val.item(), 5) == kat["integrate_theta2_phi_n10_rounded"]
    want = float.fromhex(kat["integrate_theta2_phi_n10"])
    assert abs(val.item() - want) <= INTEGRAL_REL * abs(want)
    k5 = np.asarray([float.fromhex(h) for h in kat["exp_kx_std4_k"]])
    kd = torch.from_numpy(k5).to(DEV)
    c4 = us.create_standard(4)
    f = lambda s: torch.exp(torch.einsum("v,v...->...", kd, c4.to_cartesian(s, as_array=True)))  # noqa: E731
    for n, key in ((16, "exp_kx_std4_n16"), (24, "exp_kx_std4_n24")):
        got = us.integrate(c4, f, False, n, xp=torch, device=DEV, dtype=torch.float64).item()
        assert abs(got - float.fromhex(kat[key])) <= INTEGRAL_REL * abs(got)
        assert abs(got - vo.analytic_exp_kx(k5)) <= INTEGRAL_REL * abs(got)


# ---- the reference's own tests (tests/test_integral.py) on the CUDA path ---------------------
@pytest.mark.parametrize("n", [4, 8, 16])
@pytest.mark.parametrize("value, expected", [(1, 4 * math.pi), (1j, 4j * math.pi)])
def test_sphere_surface_integrate(value, expected, n):
    def f2(s):
        return torch.asarray(value, device=DEV) * torch.ones_like(s["theta"], device=DEV)

    got = us.integrate(us.create_spherical(), f2, does_f_support_separation_of_variables=False, n=n, xp=torch, device=DEV)
    assert got.item() == pytest.approx(expected, rel=1e-2)


@pytest.mark.parametrize("r", [1, 3])
@pytest.mark.parametrize("n", [4, 8, 16])
@pytest.mark.parametrize("factory", [us.create_spherical, lambda: us.create_standard(3), lambda: us.create_hopf(2)])
@pytest.mark.parametrize("concat", [True, False])
def test_integrate_surface_area(factory, n, concat, r):
    c = factory()

    def f(s):
        if concat:
            return torch.asarray(r**c.s_ndim) * torch.ones_like(next(iter(s.values())))
        return {k: torch.asarray(r, device=DEV) * torch.ones_like(s[k], device=DEV) for k in c.s_nodes}

    actual = us.integrate(c, f, does_f_support_separation_of_variables=not concat, n=n, xp=torch, device=DEV)
    if not concat:
        actual = torch.prod(torch.stack(list(actual.values())))
    assert actual.item() == pytest.approx(c.surface_area(r), rel=1e-2)


@pytest.mark.parametrize("n", [3, 4, 5])
def test_integrate_match_across_trees(n):
    cs = [us.create_random(n - 1) for _ in range(4)]
    k = torch.rand((n,), device=DEV)

    def create_f(c):
        def f(s):
            x = c.to_cartesian(s, as_array=True)
            return torch.einsum("v,v...->...", k.to(x.dtype), x)

        return f

    actual = [us.integrate(c, create_f(c), does_f_support_separation_of_variables=False, n=8, xp=torch, device=DEV) for c in cs]
    assert torch.allclose(torch.stack(actual), actual[0], rtol=1e-3, atol=1e-3)


def test_surface_area_to_machine_precision():
    for c in (us.create_standard(4), us.create_hopf(3), us.create_standard_prime(4), us.create_random(6, rng=np.random.default_rng(1))):
        got = us.integrate(c, lambda s: torch.ones((1,) * c.s_ndim, device=DEV, dtype=torch.float64), False, 10, xp=torch, device=DEV, dtype=torch.float64)
        assert abs(got.item() - c.surface_area()) <= 1e-13 * c.surface_area()


def test_integrand_given_as_values_and_errors():
    c = us.create_standard(3)
    xs, ws = us.roots(c, 6, expand_dims_x=True, xp=torch, device=DEV, dtype=torch.float64)
    vals = torch.cos(xs["theta0"]) ** 2 * torch.sin(xs["theta1"]) ** 2 + 0 * xs["theta2"]
    a = us.integrate(c, vals, False, 6, xp=torch, device=DEV, dtype=torch.float64)
    b = us.integrate(c, lambda s: torch.cos(s["theta0"]) ** 2 * torch.sin(s["theta1"]) ** 2 + 0 * s["theta2"], False, 6, xp=torch, device=DEV, dtype=torch.float64)
    assert abs(a.item() - b.item()) <= 1e-14 * abs(a.item())
    with pytest.raises(ValueError):
        us.integrate(c, lambda s: torch.ones(3, device=DEV), False, 6, xp=torch, device=DEV)
    # float32 integrand
    got32 = us.integrate(c, lambda s: torch.cos(s["theta0"]) ** 2 + 0 * s["theta1"] + 0 * s["theta2"], False, 8, xp=torch, device=DEV, dtype=torch.float32)
    assert got32.dtype == torch.float32
    t = vo.create_standard(3)
    want = vo.integrate(t, lambda s: np.cos(s["theta0"]) ** 2 + 0 * s["theta1"] + 0 * s["theta2"], False, 8)
    assert abs(got32.item() - float(want)) <= 1e-5 * abs(float(want))


@pytest.mark.parametrize("m", [1, 2, 3, 5, 8, 32, 33, 100, 700])
def test_weighted_reduce_trailing_sizes(m):
    """Both reduction kernels (rows: m | 32, columns: any m) against float64 NumPy."""
    rng = np.random.default_rng(m)
    ext = (5, 7, 12)
    val = rng.normal(size=ext + (m,))
    ws = [rng.uniform(0.1, 1, e) for e in ext]
    want = np.einsum("i,j,k,ijkm->m", *ws, val)
    got = host(us.weighted_reduce(torch.from_numpy(val).to(DEV), [torch.from_numpy(w).to(DEV) for w in ws]))
    scale = np.einsum("i,j,k,ijkm->m", *ws, np.abs(val))
    assert np.all(np.abs(got - want) <= 1e-13 * scale)
    acc = torch.from_numpy(want.copy()).to(DEV)
    us.weighted_reduce(torch.from_numpy(val).to(DEV), [torch.from_numpy(w).to(DEV) for w in ws], out=acc)
    assert np.all(np.abs(host(acc) - 2 * want) <= 2e-13 * scale)


def test_weighted_reduce_long_inner_axis_and_single_axis():
    rng = np.random.default_rng(0)
    val = rng.normal(size=(3, 10000))
    ws = [rng.uniform(0.1, 1, 3), rng.uniform(0.1, 1, 10000)]
    want = np.einsum("i,j,ij->", *ws, val)
    got = us.weighted_reduce(torch.from_numpy(val).to(DEV), [torch.from_numpy(w).to(DEV) for w in ws]).item()
    assert abs(got - want) <= 1e-12 * np.einsum("i,j,ij->", *ws, np.abs(val))
    v1 = rng.normal(size=(100001,))
    w1 = rng.uniform(0, 1, 100001)
    got = us.weighted_reduce(torch.from_numpy(v1).to(DEV), [torch.from_numpy(w1).to(DEV)]).item()
    assert abs(got - float(w1 @ v1)) <= 1e-12 * float(w1 @ np.abs(v1))


def test_weighted_reduce_past_2_pow_31_values():
    """A 2.1e9-value integrand (8.6 GB of float32): 64-bit offsets in the TMA slab ring, the row
    weights and the two-phase final sum.  Values depend on the row so a dropped or doubled slab
    shows; the expected sum factorises."""
    free, _ = torch.cuda.mem_get_info()
    if free < 20 << 30:
        pytest.skip("needs 20 GB of free device memory")
    n0, n1, n2 = 2049, 1024, 1025  # 2049 * 1024 * 1025 = 2,150,630,400 > 2^31
    gen = torch.Generator(device=DEV).manual_seed(3)
    row = torch.rand(n0, device=DEV, dtype=torch.float32, generator=gen) + 0.5
    val = row[:, None, None].expand(n0, n1, n2).contiguous()
    ws = [torch.rand(n, device=DEV, dtype=torch.float32, generator=gen) + 0.1 for n in (n0, n1, n2)]
    got = us.weighted_reduce(val, ws).item()
    want = ((ws[0].double() * row.double()).sum() * ws[1].double().sum() * ws[2].double().sum()).item()
    assert abs(got - want) <= 1e-6 * want, (got, want)  # float32 products w0*w1*w2*val, float64 sums
    # the same as a complex, trailing-axis integrand through the column kernel
    del val
    torch.cuda.empty_cache()
    m = 40
    v2 = (row[:, None, None] * torch.ones((1, 1024 * 13, m), device=DEV)).contiguous()  # 2049*13312*40 = 1.09e9
    w2 = [ws[0], torch.rand(1024 * 13, device=DEV, generator=gen) + 0.1]
    got2 = us.weighted_reduce(v2, w2)
    want2 = ((w2[0].double() * row.double()).sum() * w2[1].double().sum()).item()
    assert got2.shape == (m,) and (got2.double() - want2).abs().max().item() <= 1e-6 * want2


def test_full_size_quadrature_matches_analytic():
    """BASELINE config 5: create_standard(4), n=96 (169,869,312 nodes), f = exp(k.x), float64."""
    k5 = np.random.default_rng(0).uniform(0, 1, 5)
    kd = torch.from_numpy(k5).to(DEV)
    c = us.create_standard(4)
    f = lambda s: torch.exp(torch.einsum("v,v...->...", kd, c.to_cartesian(s, as_array=True)))  # noqa: E731
    got = us.integrate(c, f, False, 96, xp=torch, device=DEV, dtype=torch.float64).item()
    want = vo.analytic_exp_kx(k5)
    assert abs(got - want) <= INTEGRAL_REL * want
    assert abs(got - 29.44939782558211) <= INTEGRAL_REL * want  # the reference's own n=96 value


@pytest.mark.parametrize("n", [1, 2, 3])
@pytest.mark.parametrize("spec", ["polar", "spherical", "standard:3", "hopf:2"])
def test_tiny_rules_match_oracle(spec, n):
    """Degenerate rule sizes (n = 1 gives a 1-point Gauss rule and a 2-point trapezoid)."""
    c, t = make(us, spec), vo.make(spec)
    k = np.linspace(0.3, 0.9, c.c_ndim)
    kd = torch.from_numpy(k).to(DEV)
    want = vo.integrate(t, lambda s: np.exp(np.einsum("v,v...->...", k, vo.to_cartesian(t, s, as_array=True))), False, n)
    got = us.integrate(
        c, lambda s: torch.exp(torch.einsum("v,v...->...", kd, c.to_cartesian(s, as_array=True))), False, n,
        xp=torch, device=DEV, dtype=torch.float64,
    )
    assert abs(got.item() - float(want)) <= INTEGRAL_REL * abs(float(want))
    sep_w = vo.integrate(t, lambda s: {node: np.cos(s[node]) + 2 for node in t.s_nodes}, True, n)
    sep_g = us.integrate(c, lambda s: {node: torch.cos(s[node]) + 2 for node in c.s_nodes}, True, n, xp=torch, device=DEV, dtype=torch.float64)
    for node in c.s_nodes:
        assert abs(sep_g[node].item() - float(sep_w[node])) <= 1e-13 * abs(float(sep_w[node]))


def test_integrand_variants_match_oracle():
    """Mapping given instead of a callable, float32 values under float64 rules, an integrand that
    ignores the leading angle (size-1 leading axis), and a sliced (non-contiguous) value tensor."""
    c, t = us.create_standard(3), vo.create_standard(3)
    n = 7
    xs, _ = us.roots(c, n, expand_dims_x=True, xp=torch, device=DEV, dtype=torch.float64)
    nxs, _ = vo.roots(t, n, expand_dims_x=True)
    # size-1 leading axis
    got = us.integrate(c, lambda s: torch.sin(s["theta1"]) ** 2 * torch.cos(s["theta2"]) ** 2, False, n, xp=torch, device=DEV, dtype=torch.float64)
    want = vo.integrate(t, lambda s: np.sin(s["theta1"]) ** 2 * np.cos(s["theta2"]) ** 2, False, n)
    assert got.shape == () and abs(got.item() - float(want)) <= INTEGRAL_REL * abs(float(want))
    # non-contiguous values: every other trailing component of a wider tensor
    wide = torch.stack([torch.cos(xs["theta0"]) ** 2 * (1 + 0 * xs["theta1"]) * (j + 1 + 0 * xs["theta2"]) for j in range(6)], dim=-1)
    got = us.integrate(c, wide[..., ::2], False, n, xp=torch, device=DEV, dtype=torch.float64)
    nwide = np.stack([np.cos(nxs["theta0"]) ** 2 * (1 + 0 * nxs["theta1"]) * (j + 1 + 0 * nxs["theta2"]) for j in range(6)], axis=-1)
    want = vo.integrate(t, nwide[..., ::2], False, n)
    np.testing.assert_allclose(host(got), want, rtol=INTEGRAL_REL)
    # float32 values with float64 rules promote to float64
    got = us.integrate(c, lambda s: (torch.cos(s["theta0"]) ** 2 + 0 * s["theta1"] + 0 * s["theta2"]).float(), False, n, xp=torch, device=DEV, dtype=torch.float64)
    assert got.dtype == torch.float64
    # separable mapping passed directly
    xs1, _ = us.roots(c, n, expand_dims_x=False, xp=torch, device=DEV, dtype=torch.float64)
    given = {node: torch.cos(xs1[node]) ** 2 for node in c.s_nodes}
    sep = us.integrate(c, given, True, n, xp=torch, device=DEV, dtype=torch.float64)
    nxs1, _ = vo.roots(t, n, expand_dims_x=False)
    wsep = vo.integrate(t, {node: np.cos(nxs1[node]) ** 2 for node in t.s_nodes}, True, n)
    for node in c.s_nodes:
        assert abs(sep[node].item() - float(wsep[node])) <= 1e-13 * max(1.0, abs(float(wsep[node])))


# ---- autograd ----------------------------------------------------------------------------------
def test_weighted_reduce_is_differentiable_in_the_integrand():
    gen = torch.Generator(device=DEV).manual_seed(2)
    val = torch.rand((3, 4, 5), device=DEV, dtype=torch.float64, generator=gen, requires_grad=True)
    ws = [torch.rand(3, device=DEV, dtype=torch.float64, generator=gen), torch.rand(4, device=DEV, dtype=torch.float64, generator=gen)]
    assert torch.autograd.gradcheck(lambda v: us.weighted_reduce(v, ws), (val,), eps=1e-6, atol=1e-9)
    cval = torch.randn((3, 4), device=DEV, dtype=torch.complex128, requires_grad=True)
    out = us.weighted_reduce(cval, ws)
    out.real.backward()
    assert torch.allclose(cval.grad, (ws[0][:, None] * ws[1][None, :]).to(torch.complex128))
    with pytest.raises(RuntimeError, match="weights"):
        us.weighted_reduce(val, [ws[0].clone().requires_grad_(), ws[1]])
    with torch.no_grad():
        assert not us.weighted_reduce(val, ws).requires_grad


@pytest.mark.parametrize("slab", [None, 4096])
def test_integrals_differentiate_with_respect_to_integrand_parameters(slab):
    """d/da of the integral of exp(a k.x) over S^3: autograd through the fused reduction against the
    derivative of the analytic value (central difference of the closed form)."""
    c = us.create_hopf(2)
    k = np.random.default_rng(4).uniform(0.2, 1.0, 4)
    kd = torch.from_numpy(k).to(DEV)
    a = torch.tensor(0.7, device=DEV, dtype=torch.float64, requires_grad=True)

    def f(s):
        return torch.exp(a * c.to_cartesian_dot(s, kd))

    val = us.integrate(c, f, False, 14, xp=torch, device=DEV, dtype=torch.float64, slab_bytes=slab)
    (grad,) = torch.autograd.grad(val, a)
    h = 1e-5
    want = (vo.analytic_exp_kx((0.7 + h) * k) - vo.analytic_exp_kx((0.7 - h) * k)) / (2 * h)
    assert abs(val.item() - vo.analytic_exp_kx(0.7 * k)) <= 1e-11 * abs(val.item())
    assert abs(grad.item() - want) <= 1e-8 * abs(want)
    # separable form: per-axis integrals of a^2 cos^2(theta)
    sep = us.integrate(c, lambda s: {n: a * a * torch.cos(s[n]) ** 2 for n in c.s_nodes}, True, 9, xp=torch, device=DEV, dtype=torch.float64)
    total = sum(v.sum() for v in sep.values())
    (g2,) = torch.autograd.grad(total, a)
    assert abs(g2.item() - 2 * total.item() / 0.7) <= 1e-10 * abs(g2.item())


def test_integrate_can_be_captured_in_a_cuda_graph():
    """Many small integrals are bound by host work (~140 us each); once the quadrature tables are
    cached nothing in integrate() synchronises or copies from the host, so the whole call — grid
    transform, integrand, fused reduction — can be captured and replayed with new parameters."""
    c = us.create_standard(3)
    k = torch.tensor([0.3, 0.5, 0.2, 0.7], device=DEV, dtype=torch.float64)

    def run():
        return us.integrate(c, lambda s: torch.exp(c.to_cartesian_dot(s, k)), False, 20, xp=torch, device=DEV, dtype=torch.float64)

    eager = run().item()  # also warms the table cache
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        run()
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = run()
    g.replay()
    assert abs(out.item() - eager) <= 1e-14 * abs(eager)
    k.mul_(2.0)  # new parameters, same graph
    g.replay()
    want = vo.analytic_exp_kx(2 * np.array([0.3, 0.5, 0.2, 0.7]))
    assert abs(out.item() - want) <= 1e-11 * want


def test_more_grid_axes_than_one_launch_indexes():
    """A tree with 13 angles has a 13-axis grid; the fused contraction indexes 12 axes per launch
    (US_MAX_DIMS), so the host contracts in two passes.  Checked against plain torch arithmetic."""
    k = 13
    ext = [2] * (k - 1) + [4]
    g = torch.Generator(device=DEV).manual_seed(5)
    val = torch.rand(ext + [3], device=DEV, dtype=torch.float64, generator=g)
    ws = [torch.rand(e, device=DEV, dtype=torch.float64, generator=g) for e in ext]
    got = us.weighted_reduce(val, ws)
    want = val
    for w in ws:
        want = torch.tensordot(w, want, dims=([0], [0]))
    assert got.shape == (3,)
    assert torch.allclose(got, want, rtol=1e-13, atol=0)
    c = us.create_standard(13)
    area = us.integrate(c, lambda s: torch.ones((2,) * 12 + (4,), device=DEV, dtype=torch.float64), False, 2, xp=torch, device=DEV, dtype=torch.float64)
    assert abs(area.item() - c.surface_area()) <= 0.2 * c.surface_area()  # n = 2 is a crude rule; the point is that it runs


def test_result_dtypes_follow_the_reference_promotion():
    """integrate(dtype=float32) stays float32 also when f is constant along an axis (the summed
    weight keeps the rule's dtype), and the separable branch returns each axis in promote(value,
    weight) even when other axes are wider (reference: vecdot per axis, _integral.py:245-276)."""
    c = us.create_spherical()
    one = us.integrate(c, lambda s: torch.ones((1, 1), device=DEV, dtype=torch.float32), False, 6, xp=torch, device=DEV, dtype=torch.float32)
    assert one.dtype == torch.float32 and abs(one.item() - 4 * math.pi) < 1e-5
    sep = us.integrate(c, lambda s: {k: torch.ones_like(v) for k, v in s.items()}, True, 6, xp=torch, device=DEV, dtype=None)
    xs, ws = us.roots(c, 6, expand_dims_x=False, xp=torch, device=DEV, dtype=None)
    for node in c.s_nodes:
        assert sep[node].dtype == torch.promote_types(xs[node].dtype, ws[node].dtype), node


def test_integral_graph_replays_with_updated_parameters():
    """us.IntegralGraph: the integral captured once, replayed after in-place parameter updates;
    every replay equals the eager integrate() of the same parameters."""
    c = us.create_standard(4)
    k = torch.tensor([0.3, -0.2, 0.5, 0.1, 0.7], device=DEV, dtype=torch.float64)
    f = lambda s: torch.exp(c.to_cartesian_dot(s, k))  # noqa: E731
    g = us.IntegralGraph(c, f, False, 12, xp=torch, device=DEV, dtype=torch.float64)
    for scale in (1.0, 0.5, 2.0):
        k.copy_(torch.tensor([0.3, -0.2, 0.5, 0.1, 0.7], device=DEV, dtype=torch.float64) * scale)
        got = g().clone()
        want = us.integrate(c, f, False, 12, xp=torch, device=DEV, dtype=torch.float64)
        assert torch.equal(got, want), scale
