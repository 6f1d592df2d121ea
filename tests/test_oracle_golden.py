"""The NumPy oracle against the golden vectors produced by the unmodified reference
(oracle/gen_golden.py).  CPU only."""
import json
import os

import numpy as np
import pytest

from oracle import vks_oracle as vo
from tests.helpers import GOLDEN, load_trees

TREES = load_trees()
TR = np.load(os.path.join(GOLDEN, "transforms.npz"))
QD = np.load(os.path.join(GOLDEN, "quadrature.npz"))
with open(os.path.join(GOLDEN, "kat.json")) as fh:
    KAT = json.load(fh)

TRANSFORM_SPECS = sorted({k.split("|")[0] for k in TR.files})
QUAD_SPECS = sorted({k.split("|")[0] for k in QD.files})


@pytest.mark.parametrize("spec", sorted(TREES))
def test_tree_structure(spec):
    t = vo.make(spec)
    g = TREES[spec]
    assert t.s_nodes == g["s_nodes"]
    assert t.c_nodes == g["c_nodes"]
    assert [list(e) for e in t.cos_edges] == g["cos_edges"]
    assert [list(e) for e in t.sin_edges] == g["sin_edges"]
    assert [[k, v] for k, v in t.branching_types.items()] == g["branching_types"]
    assert [[k, (None if v != v else int(v))] for k, v in t.S.items()] == g["S"]
    assert t.root == g["root"]
    assert (t.s_ndim, t.c_ndim) == (g["s_ndim"], g["c_ndim"])
    if t.s_ndim > 0:
        assert t.expression_str == g["expression_str"]


def _close_ulp(a, b, ulps):
    d = vo.ulp_distance(a, b)
    return float(np.max(d)) <= ulps


@pytest.mark.parametrize("dname", ["f64", "f32"])
@pytest.mark.parametrize("spec", TRANSFORM_SPECS)
def test_transforms(spec, dname):
    t = vo.make(spec)
    p = f"{spec}|{dname}|"
    x = TR[p + "x"]
    sph = vo.from_cartesian(t, x)
    assert list(sph) == ["r"] + list(reversed(t.s_nodes))
    got = np.stack([sph[k] for k in sph])
    want = TR[p + "sph"]
    assert got.dtype == want.dtype
    # r is sqrt/add/mul only: bit-exact on every host
    assert np.array_equal(got[0], want[0])
    # atan2 / sin / cos are SIMD-dispatched in NumPy: allow 2 ulp between hosts
    assert _close_ulp(got[1:], want[1:], 2)
    cart = vo.to_cartesian(t, {k: want[i] for i, k in enumerate(sph)}, as_array=True)
    assert _close_ulp(cart, TR[p + "cart"], 2 * max(1, t.s_ndim))
    ang = TR[p + "ang"]
    free = {k: ang[i] for i, k in enumerate(t.s_nodes)}
    assert _close_ulp(vo.to_cartesian(t, dict(free), as_array=True), TR[p + "cart_free_no_r"], 2 * max(1, t.s_ndim))
    free["r"] = TR[p + "r"]
    assert _close_ulp(vo.to_cartesian(t, free, as_array=True), TR[p + "cart_free"], 2 * max(1, t.s_ndim))
    with np.errstate(all="ignore"):
        esph = vo.from_cartesian(t, TR[p + "edge_x"])
    egot = np.stack([esph[k] for k in esph])
    assert np.array_equal(egot[0], TR[p + "edge_sph"][0], equal_nan=True)
    assert _close_ulp(egot[1:], TR[p + "edge_sph"][1:], 2)


@pytest.mark.parametrize("spec", QUAD_SPECS)
def test_quadrature(spec):
    t = vo.make(spec)
    k = QD[f"{spec}|k"]
    for n in (4, 7, 12):
        xs, ws = vo.roots(t, n, expand_dims_x=False)
        for i, node in enumerate(t.s_nodes):
            # Golub–Welsch in LAPACK: allow last-bit differences between hosts
            np.testing.assert_allclose(xs[node], QD[f"{spec}|n{n}|x{i}"], rtol=1e-14, atol=1e-15)
            np.testing.assert_allclose(ws[node], QD[f"{spec}|n{n}|w{i}"], rtol=1e-13, atol=1e-300)
        if f"{spec}|n{n}|dense" not in QD.files:
            continue

        def f(s):
            return np.exp(np.einsum("v,v...->...", k, vo.to_cartesian(t, s, as_array=True)))

        def fv(s):
            x = vo.to_cartesian(t, s, as_array=True)
            ph = np.einsum("v,v...->...", k, x)
            return np.stack([np.exp(1j * ph), ph**2 + 0j, np.exp(ph) + 1j * x[0]], axis=-1)

        def fsep(s):
            return {
                node: np.stack([np.cos(s[node]) ** 2, 1.0 + 0 * s[node], np.sin(s[node])], axis=-1)
                for node in t.s_nodes
            }

        np.testing.assert_allclose(vo.integrate(t, f, False, n), QD[f"{spec}|n{n}|dense"], rtol=1e-13)
        np.testing.assert_allclose(vo.integrate(t, fv, False, n), QD[f"{spec}|n{n}|dense_vec"], rtol=1e-12, atol=1e-13)
        sep = vo.integrate(t, fsep, True, n)
        np.testing.assert_allclose(
            np.stack([sep[node] for node in t.s_nodes]), QD[f"{spec}|n{n}|sep"], rtol=1e-13, atol=1e-14
        )


def test_known_answers():
    c3 = vo.create_spherical()
    s = vo.from_cartesian(c3, np.asarray([1.0, 2.0, 3.0]))
    for key, hexval in KAT["spherical_from_123"].items():
        assert abs(float(s[key]) - float.fromhex(hexval)) <= 2 * np.spacing(abs(float.fromhex(hexval)))
    for key, val in KAT["spherical_from_123_rounded"].items():
        assert round(float(s[key]), 6) == val  # the reference's own doctest pin
    back = vo.to_cartesian(c3, s)
    for key, hexval in KAT["spherical_to_back"].items():
        assert abs(float(back[int(key)]) - float.fromhex(hexval)) <= 4 * np.spacing(abs(float.fromhex(hexval)))
    val = vo.integrate(c3, lambda s: s["theta"] ** 2 * s["phi"], False, 10)
    assert round(float(val), 5) == KAT["integrate_theta2_phi_n10_rounded"]
    assert abs(float(val) - float.fromhex(KAT["integrate_theta2_phi_n10"])) < 1e-12 * abs(float(val))
    for name, factory, x in (
        ("hopf2", lambda: vo.create_hopf(2), KAT["hopf2_x"]),
        ("standard3", lambda: vo.create_standard(3), KAT["hopf2_x"]),
        ("standard_prime2", lambda: vo.create_standard_prime(2), KAT["standard_prime2_x"]),
    ):
        got = vo.from_cartesian(factory(), np.asarray(x, dtype=np.float64))
        for key, hexval in KAT[name].items():
            k = key if key == "r" else key
            want = float.fromhex(hexval)
            assert abs(float(got[k]) - want) <= 2 * np.spacing(abs(want)), (name, key)
        assert float(got["r"]).hex() == KAT[name]["r"]
    # analytic value of the exp(k.x) integral on S^4
    k5 = np.asarray([float.fromhex(h) for h in KAT["exp_kx_std4_k"]])
    c4 = vo.create_standard(4)
    got = vo.integrate(
        c4, lambda s: np.exp(np.einsum("v,v...->...", k5, vo.to_cartesian(c4, s, as_array=True))), False, 16
    )
    assert abs(float(got) - vo.analytic_exp_kx(k5)) < 1e-12 * abs(float(got))
    assert abs(float(got) - float.fromhex(KAT["exp_kx_std4_n16"])) < 1e-13 * abs(float(got))
    r5 = KAT["roots_spherical_n5"]
    xs, ws = vo.roots(c3, 5, expand_dims_x=False)
    np.testing.assert_allclose(xs["theta"], [float.fromhex(h) for h in r5["x_theta"]], rtol=1e-14)
    np.testing.assert_allclose(ws["theta"], [float.fromhex(h) for h in r5["w_theta"]], rtol=1e-13)
    np.testing.assert_allclose(xs["phi"], [float.fromhex(h) for h in r5["x_phi"]], rtol=1e-15)
    np.testing.assert_allclose(ws["phi"], [float.fromhex(h) for h in r5["w_phi"]], rtol=1e-15)


def test_ulp_distance():
    a = np.asarray([1.0, -1.0, 0.0, -0.0, np.nan, 1.0])
    b = np.asarray([np.nextafter(1.0, 2.0), np.nextafter(-1.0, -2.0), -0.0, 5e-324, np.nan, np.nan])
    d = vo.ulp_distance(a, b)
    assert list(d[:5]) == [1.0, 1.0, 0.0, 1.0, 0.0] and np.isinf(d[5])
    f = np.float32
    assert vo.ulp_distance(np.asarray([f(1.0)]), np.asarray([np.nextafter(f(1.0), f(2.0))]))[0] == 1.0
