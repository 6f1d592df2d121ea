"""Shared test helpers: tree specs, a NumPy emulator of the compiled program, tolerances."""
from __future__ import annotations

import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# Parity tolerances (BASELINE.json north_star): transforms and round trips must match the
# reference's NumPy path to <= 4 ulp in float64 and <= 2e-6 relative in float32; integrals to
# <= 1e-12 relative in float64.  "Relative" is element-wise, with the magnitude of the point
# (r, which every leaf is a fraction of) as the floor of the denominator for leaves, and pi for
# angles; see DESIGN.md "Tolerances".
F64_ULP = 4
F32_REL = 2e-6
INTEGRAL_REL = 1e-12


def load_trees() -> dict:
    with open(os.path.join(GOLDEN, "trees.json")) as fh:
        return json.load(fh)


def make(us, spec: str):
    """Build the coordinate object for a spec string with any package exposing create_*."""
    head, *args = spec.split(":")
    if head == "random":
        return us.create_random(int(args[0]), rng=np.random.default_rng(int(args[1])))
    if head == "expr":
        return us.create_from_branching_types(args[0] if args else "")
    if head in ("polar", "spherical"):
        return getattr(us, f"create_{head}")()
    return getattr(us, f"create_{head}")(int(args[0]))


def emulate_to_cartesian(ct, angles: list[np.ndarray], r: np.ndarray | None) -> list[np.ndarray]:
    """NumPy rendition of the kernel's forward walk over the compiled program."""
    dt = angles[0].dtype
    cur = np.ones_like(angles[0]) if r is None else r.astype(dt)
    stack: list[np.ndarray] = []
    out: list[np.ndarray | None] = [None] * ct.n_leaves
    for k in range(ct.n_nodes):
        th = angles[ct.angle_slot[k]]
        pc = cur * np.cos(th)
        ps = cur * np.sin(th)
        kind = ct.kinds[k]
        if kind == 0:
            out[ct.cos_leaf[k]] = pc
            out[ct.sin_leaf[k]] = ps
            if stack:
                cur = stack.pop()
        elif kind == 1:
            out[ct.cos_leaf[k]] = pc
            cur = ps
        elif kind == 2:
            out[ct.sin_leaf[k]] = ps
            cur = pc
        else:
            stack.append(ps)
            cur = pc
    assert not stack
    return out  # type: ignore[return-value]


def emulate_from_cartesian(ct, x: list[np.ndarray]) -> tuple[np.ndarray, list[np.ndarray]]:
    """NumPy rendition of the kernel's backward walk (children before parents)."""
    acc = x[0] * x[0]
    for l in range(1, ct.n_leaves):
        acc = acc + x[l] * x[l]
    r = np.sqrt(acc)
    stack: list[np.ndarray] = []
    cur = None
    angles: list[np.ndarray | None] = [None] * ct.n_nodes
    for k in range(ct.n_nodes - 1, -1, -1):
        kind = ct.kinds[k]
        if kind == 0:
            if k != ct.n_nodes - 1:
                stack.append(cur)
            vc, vs = x[ct.cos_leaf[k]], x[ct.sin_leaf[k]]
        elif kind == 1:
            vc, vs = x[ct.cos_leaf[k]], cur
        elif kind == 2:
            vc, vs = cur, x[ct.sin_leaf[k]]
        else:
            vc, vs = cur, stack.pop()
        angles[ct.angle_slot[k]] = np.arctan2(vs, vc)
        cur = np.sqrt(vs * vs + vc * vc)
    assert not stack
    return r, angles  # type: ignore[return-value]
