"""Generate tests/golden/* from the UNMODIFIED reference (build container only).

    python -m oracle.gen_golden

Imports /root/reference through oracle/refload.py (NumPy backend), evaluates the reference's
own ``create_*``, ``from_cartesian``, ``to_cartesian``, ``roots`` and ``integrate`` on seeded
inputs and stores inputs + outputs:

  tests/golden/trees.json        structure of every tree spec (orders, edges, S, types)
  tests/golden/transforms.npz    per spec × dtype: x → from_cartesian → to_cartesian, free angles
                                 → to_cartesian, and the edge vectors of SURVEY.md §8d
  tests/golden/quadrature.npz    roots tables, dense integrals of exp(k·x), separable integrals
  tests/golden/kat.json          the numeric pins of the reference's own doctests/README plus the
                                 hex known-answers listed in SURVEY.md §8c

NumPy's transcendental ufuncs are SIMD-dispatched per host CPU, so consumers compare the
sin/cos/atan2-derived arrays within 2 ulp and the sqrt/add-derived ones (``r``) bit-for-bit.
"""
from __future__ import annotations

import json
import os
import platform

import numpy as np
import scipy

from oracle import refload

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")

TREE_SPECS = [
    "polar",
    "spherical",
    "standard:0",
    "standard:1",
    "standard:2",
    "standard:3",
    "standard:4",
    "standard:9",
    "standard:40",
    "standard_prime:0",
    "standard_prime:1",
    "standard_prime:3",
    "standard_prime:8",
    "hopf:0",
    "hopf:1",
    "hopf:2",
    "hopf:3",
    "hopf:5",
    "random:1:0",
    "random:5:0",
    "random:6:3",
    "random:32:0",
    "random:32:1",
    "random:103:7",
    "expr:cbaa",
    "expr:cbbaa",
    "expr:cbabba",
    "expr:cab'a",
    "expr:bpbpa",
]
TRANSFORM_SPECS = [
    s
    for s in TREE_SPECS
    if s not in ("standard:0", "standard_prime:0", "hopf:0", "standard:40", "random:103:7", "hopf:5")
]
QUAD_SPECS = [
    "polar",
    "spherical",
    "standard:3",
    "standard:4",
    "standard_prime:3",
    "hopf:2",
    "hopf:3",
    "random:5:0",
    "random:6:3",
    "expr:cbaa",
    "expr:cbabba",
]
N_POINTS = 192


def make(us, spec: str):
    head, *args = spec.split(":")
    if head == "random":
        return us.create_random(int(args[0]), rng=np.random.default_rng(int(args[1])))
    if head == "expr":
        return us.create_from_branching_types(args[0] if args else "")
    if head in ("polar", "spherical"):
        return getattr(us, f"create_{head}")()
    return getattr(us, f"create_{head}")(int(args[0]))


def edge_vectors(c_ndim: int, dtype) -> np.ndarray:
    """zeros, signed zeros, ±axis points, equal components, huge/tiny magnitudes."""
    cols = [np.zeros(c_ndim), -np.zeros(c_ndim) * 1.0]
    cols[1] = np.full(c_ndim, -0.0)
    for k in range(min(c_ndim, 6)):
        e = np.zeros(c_ndim)
        e[k] = 1.0
        cols.append(e)
        cols.append(-e)
        m = np.full(c_ndim, -0.0)
        m[k] = -2.5
        cols.append(m)
    cols.append(np.ones(c_ndim))
    cols.append(-np.ones(c_ndim))
    cols.append(np.full(c_ndim, 1e30))
    cols.append(np.full(c_ndim, 1e-30))
    alt = np.where(np.arange(c_ndim) % 2 == 0, 3e-20, -7e18)
    cols.append(alt)
    return np.stack(cols, axis=1).astype(dtype)


def main() -> None:
    R = refload.load()
    us = R.us
    os.makedirs(OUT, exist_ok=True)

    # ---- trees -----------------------------------------------------------------------
    trees = {}
    for spec in TREE_SPECS:
        c = make(us, spec)
        trees[spec] = {
            "s_nodes": list(c.s_nodes),
            "c_nodes": list(c.c_nodes),
            "cos_edges": [list(e) for e in c.cos_edges],
            "sin_edges": [list(e) for e in c.sin_edges],
            "branching_types": [[k, v.value] for k, v in c.branching_types.items()],
            "S": [[k, (None if v != v else int(v))] for k, v in c.S.items()],
            "root": c.root,
            "s_ndim": c.s_ndim,
            "c_ndim": c.c_ndim,
            "expression_str": c.branching_types_expression_str if c.s_ndim > 0 else None,
            "repr": repr(c) if c.s_ndim > 0 else None,
            "G_nodes": list(c.G.nodes),
        }
    with open(os.path.join(OUT, "trees.json"), "w") as fh:
        json.dump(trees, fh, indent=1, sort_keys=True)

    # ---- transforms --------------------------------------------------------------------
    tr = {}
    for spec in TRANSFORM_SPECS:
        c = make(us, spec)
        for dname, dt in (("f64", np.float64), ("f32", np.float32)):
            rng = np.random.default_rng(abs(hash_spec(spec)) % (2**32))
            x = rng.uniform(-1, 1, (c.c_ndim, N_POINTS)).astype(dt)
            sph = c.from_cartesian(x)  # array input: indexed by leaf id (README.md:68)
            key_order = list(sph.keys())
            assert key_order == ["r"] + list(reversed(c.s_nodes))
            sph_arr = np.stack([sph[k] for k in key_order])
            cart = c.to_cartesian(sph, as_array=True)
            # free angles well outside (-pi, pi] to pin the range reduction, r in (0, 2)
            ang = rng.uniform(-20.0, 20.0, (c.s_ndim, N_POINTS)).astype(dt)
            rr = rng.uniform(0.0, 2.0, N_POINTS).astype(dt)
            free = {k: ang[i] for i, k in enumerate(c.s_nodes)} | {"r": rr}
            cart2 = c.to_cartesian(free, as_array=True)
            free_no_r = {k: ang[i] for i, k in enumerate(c.s_nodes)}
            cart3 = c.to_cartesian(free_no_r, as_array=True)
            ex = edge_vectors(c.c_ndim, dt)
            with np.errstate(all="ignore"):
                esph = c.from_cartesian(ex)
                esph_arr = np.stack([esph[k] for k in key_order])
                ecart = c.to_cartesian(esph, as_array=True)
            p = f"{spec}|{dname}|"
            tr[p + "x"] = x
            tr[p + "sph"] = sph_arr
            tr[p + "cart"] = cart
            tr[p + "ang"] = ang
            tr[p + "r"] = rr
            tr[p + "cart_free"] = cart2
            tr[p + "cart_free_no_r"] = cart3
            tr[p + "edge_x"] = ex
            tr[p + "edge_sph"] = esph_arr
            tr[p + "edge_cart"] = ecart
    np.savez_compressed(os.path.join(OUT, "transforms.npz"), **tr)

    # ---- quadrature --------------------------------------------------------------------
    qd = {}
    for spec in QUAD_SPECS:
        c = make(us, spec)
        rng = np.random.default_rng(abs(hash_spec(spec)) % (2**32) + 1)
        k = rng.uniform(0, 1, c.c_ndim)
        qd[f"{spec}|k"] = k
        for n in (4, 7, 12):
            xs, ws = us.roots(c, n, expand_dims_x=False, xp=np)
            for i, node in enumerate(c.s_nodes):
                qd[f"{spec}|n{n}|x{i}"] = xs[node]
                qd[f"{spec}|n{n}|w{i}"] = ws[node]
            if n ** c.s_ndim > 5e6:
                continue

            def f(s, c=c, k=k):
                x = c.to_cartesian(s, as_array=True)
                return np.exp(np.einsum("v,v...->...", k, x))

            def fv(s, c=c, k=k):  # vector-valued complex integrand: trailing axis of 3
                x = c.to_cartesian(s, as_array=True)
                ph = np.einsum("v,v...->...", k, x)
                return np.stack([np.exp(1j * ph), ph**2 + 0j, np.exp(ph) + 1j * x[0]], axis=-1)

            def fsep(s, c=c):
                return {
                    node: np.stack([np.cos(s[node]) ** 2, 1.0 + 0 * s[node], np.sin(s[node])], axis=-1)
                    for node in c.s_nodes
                }

            qd[f"{spec}|n{n}|dense"] = np.asarray(us.integrate(c, f, False, n, xp=np))
            qd[f"{spec}|n{n}|dense_vec"] = np.asarray(us.integrate(c, fv, False, n, xp=np))
            sep = us.integrate(c, fsep, True, n, xp=np)
            qd[f"{spec}|n{n}|sep"] = np.stack([sep[node] for node in c.s_nodes])
    np.savez_compressed(os.path.join(OUT, "quadrature.npz"), **qd)

    # ---- known answers -----------------------------------------------------------------
    c3 = us.create_spherical()
    s3 = c3.from_cartesian(np.asarray([1.0, 2.0, 3.0]))
    back = c3.to_cartesian(s3)
    k5 = np.random.default_rng(0).uniform(0, 1, 5)
    c4 = us.create_standard(4)

    def f5(s):
        return np.exp(np.einsum("v,v...->...", k5, c4.to_cartesian(s, as_array=True)))

    def hexdict(c, x):
        s = c.from_cartesian(np.asarray(x, dtype=np.float64))
        return {str(k): float(v).hex() for k, v in s.items()}

    kat = {
        "_provenance": {
            "generator": "oracle/gen_golden.py",
            "reference": "ultrasphere 2.0.4 (/root/reference), NumPy backend",
            "numpy": np.__version__,
            "scipy": scipy.__version__,
            "machine": platform.machine(),
        },
        # reference doctests: _coordinates.py:545-552, :627-634
        "spherical_from_123": {str(k): float(v).hex() for k, v in s3.items()},
        "spherical_from_123_rounded": {"theta": 0.640522, "phi": 1.107149, "r": 3.741657},
        "spherical_to_back": {str(k): float(v).hex() for k, v in back.items()},
        # _integral.py:212-225 / README.md:156-167 (rounded to 5 places there: 110.02621)
        "integrate_theta2_phi_n10": float(
            us.integrate(c3, lambda s: s["theta"] ** 2 * s["phi"], False, 10, xp=np)
        ).hex(),
        "integrate_theta2_phi_n10_rounded": 110.02621,
        # SURVEY.md §8c hex KATs
        "hopf2_x": [0.25, -0.5, 0.75, -1.0],
        "hopf2": hexdict(us.create_hopf(2), [0.25, -0.5, 0.75, -1.0]),
        "standard3": hexdict(us.create_standard(3), [0.25, -0.5, 0.75, -1.0]),
        "standard_prime2_x": [1 / 3, -2 / 3, 1.0],
        "standard_prime2": hexdict(us.create_standard_prime(2), [1 / 3, -2 / 3, 1.0]),
        # analytic KAT: create_standard(4), k = default_rng(0).uniform(0,1,5)
        "exp_kx_std4_k": [float(v).hex() for v in k5],
        "exp_kx_std4_n16": float(us.integrate(c4, f5, False, 16, xp=np)).hex(),
        "exp_kx_std4_n24": float(us.integrate(c4, f5, False, 24, xp=np)).hex(),
        # _integral.py:59-78 doctest table (5-point rule on `ba`), full precision
        "roots_spherical_n5": {
            "x_theta": [float(v).hex() for v in us.roots(c3, 5, expand_dims_x=False, xp=np)[0]["theta"]],
            "w_theta": [float(v).hex() for v in us.roots(c3, 5, expand_dims_x=False, xp=np)[1]["theta"]],
            "x_phi": [float(v).hex() for v in us.roots(c3, 5, expand_dims_x=False, xp=np)[0]["phi"]],
            "w_phi": [float(v).hex() for v in us.roots(c3, 5, expand_dims_x=False, xp=np)[1]["phi"]],
        },
        "surface_area": {str(d): float(us.create_standard(d - 1).surface_area()) for d in (2, 3, 4, 5, 8, 10)},
    }
    with open(os.path.join(OUT, "kat.json"), "w") as fh:
        json.dump(kat, fh, indent=1, sort_keys=True)
    sizes = {f: os.path.getsize(os.path.join(OUT, f)) for f in sorted(os.listdir(OUT))}
    print("wrote", sizes)


def hash_spec(spec: str) -> int:
    """Stable (process-independent) seed from a spec string."""
    h = 2166136261
    for ch in spec.encode():
        h = ((h ^ ch) * 16777619) % (2**32)
    return h


if __name__ == "__main__":
    main()
