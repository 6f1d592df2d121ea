"""Seeded fuzzing of the CUDA transforms against the NumPy oracle: random trees, random batch
shapes, per-operand broadcasting, non-contiguous views, optional / scalar radius — the input
generality the reference gets for free from array broadcasting (_coordinates.py:636-656, 555-579)."""
import numpy as np
import pytest
import torch

import ultrasphere_b200 as us
from oracle import vks_oracle as vo
from tests.test_gpu_parity import assert_angles_close, assert_leaves_close, depth_of, dev, host

pytestmark = pytest.mark.gpu
KINDS = ["a", "b", "b'", "c"]


def random_expression(rng: np.random.Generator, s_ndim: int) -> str:
    """A valid branching expression with s_ndim internal nodes: grow a binary tree by splitting
    random leaves, then list it in pre-order (cos subtree first)."""
    kids = {0: None}  # node -> (cos, sin) or None for a leaf
    leaves = [0]
    for _ in range(s_ndim):
        at = leaves.pop(int(rng.integers(len(leaves))))
        a, b = len(kids), len(kids) + 1
        kids[at] = (a, b)
        kids[a] = kids[b] = None
        leaves += [a, b]
    out = []

    def walk(n):
        cs = kids[n]
        if cs is None:
            return
        cos_leaf, sin_leaf = kids[cs[0]] is None, kids[cs[1]] is None
        out.append("a" if cos_leaf and sin_leaf else "b" if cos_leaf else "b'" if sin_leaf else "c")
        walk(cs[0])
        walk(cs[1])

    walk(0)
    return "".join(out)


def operand(rng, shape, ndt, lo, hi):
    """Random array broadcastable to `shape`, returned as (numpy value, device tensor) where the
    device tensor may be a non-contiguous or expanded view with the same values."""
    own = [d if rng.random() < 0.6 else 1 for d in shape]
    drop = int(rng.integers(0, len(shape) + 1)) if rng.random() < 0.3 else 0
    while drop and own[:drop] != [1] * drop:
        drop -= 1
    own = own[drop:]  # fewer leading dims, like NumPy broadcasting allows
    value = rng.uniform(lo, hi, own).astype(ndt)
    mode = rng.integers(4)
    if mode == 0 or value.ndim == 0 or value.size == 0:
        return value, dev(value)
    if mode == 1:  # every other element of a wider buffer along the last axis
        wide = np.repeat(value, 2, axis=-1)
        return value, dev(wide)[..., ::2]
    if mode == 2 and value.ndim >= 2:  # transposed storage
        return value, dev(np.ascontiguousarray(np.swapaxes(value, -1, -2))).transpose(-1, -2)
    if mode == 3 and value.shape[0] > 1:  # stride-0 leading axis
        value = np.broadcast_to(value[:1], value.shape).copy()
        return value, dev(value[:1]).expand(*value.shape)
    return value, dev(value)


@pytest.mark.parametrize("seed", range(40))
def test_random_trees_shapes_and_layouts(seed):
    rng = np.random.default_rng(1000 + seed)
    s_ndim = int(rng.choice([1, 2, 3, 5, 8, 13, 21, 34, 70]))
    expr = random_expression(rng, s_ndim)
    c, t = us.create_from_branching_types(expr), vo.tree_from_branching(expr)
    assert c.s_ndim == s_ndim and list(map(str, c.s_nodes)) == list(map(str, t.s_nodes))
    ndt, tdt = (np.float32, torch.float32) if rng.random() < 0.5 else (np.float64, torch.float64)
    ndim = int(rng.integers(1, 4))
    shape = [int(rng.choice([1, 2, 3, 5, 17, 64, 257])) for _ in range(ndim)]
    if rng.random() < 0.25:
        shape[-1] = int(rng.choice([4096, 70001]))  # long enough for the tiled / ring kernels
    d = depth_of(c)

    # ---- to_cartesian: every angle (and r) with its own broadcastable shape and layout
    ang_np, ang_dev = {}, {}
    for node in c.s_nodes:
        ang_np[node], ang_dev[node] = operand(rng, shape, ndt, -3.2, 3.2)
    r_mode = rng.integers(4)
    if r_mode == 1:
        ang_np["r"] = ang_dev["r"] = 1.75
    elif r_mode == 2:
        ang_np["r"], ang_dev["r"] = operand(rng, shape, ndt, 0.25, 4.0)
    elif r_mode == 3:
        ang_np["r"] = np.asarray(0.5, dtype=ndt)
        ang_dev["r"] = torch.tensor(0.5, dtype=tdt, device="cuda:0")
    want = vo.to_cartesian(t, dict(ang_np))
    got = c.to_cartesian(dict(ang_dev))
    assert list(got) == list(want)
    r_full = np.asarray(ang_np.get("r", 1.0), dtype=np.float64)
    for i, leaf in enumerate(c.c_nodes):
        assert tuple(got[leaf].shape) == np.shape(want[leaf]), (expr, leaf)
        w = np.asarray(want[leaf], dtype=ndt)
        rr = np.broadcast_to(r_full, np.broadcast_shapes(r_full.shape, w.shape)) if r_full.ndim <= w.ndim else np.abs(r_full).max()
        if np.shape(rr) != w.shape:
            rr = np.full(w.shape, np.abs(r_full).max())
        assert_leaves_close(host(got[leaf])[None], w[None], rr, d[i : i + 1], f"{expr} leaf {leaf}")
    wantf = vo.to_cartesian(t, dict(ang_np), as_array=True).astype(ndt)
    gotf = c.to_cartesian(dict(ang_dev), as_array=True)
    assert tuple(gotf.shape) == wantf.shape and gotf.dtype == tdt
    rr = np.broadcast_to(r_full, wantf.shape[1:]) if r_full.ndim <= wantf.ndim - 1 else r_full
    assert_leaves_close(host(gotf), wantf, rr, d, expr)

    # ---- from_cartesian: leaves with their own shapes and layouts, mapping input
    x_np, x_dev = {}, {}
    for leaf in c.c_nodes:
        x_np[leaf], x_dev[leaf] = operand(rng, shape, ndt, -1.0, 1.0)
    wants = vo.from_cartesian(t, dict(x_np))
    gots = c.from_cartesian(dict(x_dev))
    assert list(map(str, gots)) == list(map(str, wants))
    for k in wants:
        assert tuple(gots[k].shape) == np.shape(wants[k]), (expr, k)
        if k == "r":
            assert np.array_equal(host(gots[k]), np.asarray(wants[k], dtype=ndt)), expr
        else:
            assert_angles_close(host(gots[k]), np.asarray(wants[k], dtype=ndt), f"{expr} {k}")

    # ---- array input (c_ndim, *shape), possibly a strided view, and the round trip
    full = rng.uniform(-1, 1, [c.c_ndim] + shape + [2]).astype(ndt)
    xa = dev(full)[..., 1]
    wa = vo.from_cartesian(t, full[..., 1])
    ga = c.from_cartesian(xa)
    assert np.array_equal(host(ga["r"]), wa["r"])
    for k in c.s_nodes:
        assert_angles_close(host(ga[k]), wa[k], f"{expr} {k} (array)")
    back = c.to_cartesian(ga, as_array=True)
    tol = 64 * np.finfo(ndt).eps * (1 + d.max())
    assert np.abs(host(back) - full[..., 1]).max() <= tol * max(1.0, float(np.abs(wa["r"]).max()))
