"""Tree compiler: ``SphericalCoordinates`` → ``us_plan_t`` (the pre-order program the kernels run).

The reference walks the networkx graph on every transform call (``nx.topological_sort`` and
``get_child`` per node — /root/reference/src/ultrasphere/_coordinates.py:566-574, 638-644; 0.1–1 ms
per call).  Here the walk happens once per coordinate object: the tree is listed in pre-order
with the cos subtree first (the order of ``branching_types_expression``, _coordinates.py:206-242),
each node recording its kind, the position of its angle in ``s_nodes`` and the positions of its
leaf children in ``c_nodes``.  ``us_plan_compile`` validates and packs that listing; the result is
cached on the object and passed by value into every launch.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Any

from . import _native
from ._coordinates import BranchingType, SphericalCoordinates, get_child

_KIND = {BranchingType.A: 0, BranchingType.B: 1, BranchingType.BP: 2, BranchingType.C: 3}


@dataclass(frozen=True)
class CompiledTree:
    """Host-side view of a compiled tree (everything index-based, no graph objects), plus what the
    dense fast paths of ``_transform`` would otherwise rebuild on every call: the plan reference,
    the bound entry points, a reusable pointer table and the output dict order."""

    plan: _native.UsPlan
    n_nodes: int
    n_leaves: int
    max_stack: int
    kinds: tuple[int, ...]          # per program position
    angle_slot: tuple[int, ...]     # per program position → index in s_nodes
    cos_leaf: tuple[int, ...]       # per program position → index in c_nodes or -1
    sin_leaf: tuple[int, ...]
    # per s_nodes index: indices (into s_nodes) of the node itself and all its ancestors, and
    # per c_nodes index: indices (into s_nodes) of the ancestors — used for output shapes
    angle_ancestors: tuple[tuple[int, ...], ...]
    leaf_ancestors: tuple[tuple[int, ...], ...]
    # per s_nodes index: indices (into c_nodes) of the leaves below the node
    angle_leaves: tuple[tuple[int, ...], ...]
    plan_ref: Any = None        # ctypes.byref(plan)
    from_rows: Any = None       # lib.us_from_cartesian_rows
    to_rows: Any = None         # lib.us_to_cartesian_rows
    angle_ptrs: Any = None      # (c_void_p * n_nodes)()
    ones_leaves: Any = None     # (c_uint32 * n_leaves)(1, ...): "spans the launch" masks
    from_order: tuple[tuple[Any, int], ...] = ()  # (node, row of from_cartesian's output) in dict order


def preorder_listing(c: SphericalCoordinates[Any, Any]) -> tuple[list[Any], list[int], list[Any]]:
    """Nodes in program order, their kinds, and the leaf operands in program order."""
    G = c.G
    nodes: list[Any] = []
    kinds: list[int] = []
    leaves: list[Any] = []
    pending = [c.root]
    while pending:
        node = pending.pop()
        cos_child = get_child(G, node, "cos")
        sin_child = get_child(G, node, "sin")
        kind = c.branching_types[node]
        nodes.append(node)
        kinds.append(_KIND[kind])
        if kind in (BranchingType.A, BranchingType.B):
            leaves.append(cos_child)
        if kind in (BranchingType.A, BranchingType.BP):
            leaves.append(sin_child)
        if kind in (BranchingType.B, BranchingType.C):
            pending.append(sin_child)
        if kind in (BranchingType.BP, BranchingType.C):
            pending.append(cos_child)
    return nodes, kinds, leaves


def compile_tree(c: SphericalCoordinates[Any, Any]) -> CompiledTree:
    if c.s_ndim < 1:
        raise ValueError("a tree without angles has nothing to compile")
    if c.s_ndim > _native.US_MAX_NODES:
        raise ValueError(
            f"tree has {c.s_ndim} angles; the CUDA plan holds at most {_native.US_MAX_NODES}"
        )
    nodes, kinds, leaves = preorder_listing(c)
    s_index = {node: i for i, node in enumerate(c.s_nodes)}
    c_index = {node: i for i, node in enumerate(c.c_nodes)}
    n = len(nodes)
    kinds_arr = (C.c_uint8 * n)(*kinds)
    slot_arr = (C.c_int32 * n)(*[s_index[node] for node in nodes])
    leaf_arr = (C.c_int32 * (n + 1))(*[c_index[leaf] for leaf in leaves])
    plan = _native.UsPlan()
    _native.check(
        _native.lib().us_plan_compile(n, kinds_arr, slot_arr, leaf_arr, C.byref(plan)), "us_plan_compile"
    )
    cos_leaf: list[int] = []
    sin_leaf: list[int] = []
    for k in range(n):
        w = plan.node[k]
        cl, sl = (w >> 24) & 0xFFFF, (w >> 40) & 0xFFFF
        cos_leaf.append(-1 if cl == 0xFFFF else cl)
        sin_leaf.append(-1 if sl == 0xFFFF else sl)
    # ancestor tables
    parent: dict[Any, Any] = {}
    for node in c.s_nodes:
        for kind in ("cos", "sin"):
            parent[get_child(c.G, node, kind)] = node

    def chain(node: Any) -> tuple[int, ...]:
        out = []
        while node in parent:
            node = parent[node]
            out.append(s_index[node])
        return tuple(out)

    angle_anc = tuple((s_index[node],) + chain(node) for node in c.s_nodes)
    leaf_anc = tuple(chain(leaf) for leaf in c.c_nodes)
    below: list[list[int]] = [[] for _ in c.s_nodes]
    for li, anc in enumerate(leaf_anc):
        for a in anc:
            below[a].append(li)
    return CompiledTree(
        plan=plan,
        n_nodes=n,
        n_leaves=n + 1,
        max_stack=int(plan.max_stack),
        kinds=tuple(kinds),
        angle_slot=tuple(s_index[node] for node in nodes),
        cos_leaf=tuple(cos_leaf),
        sin_leaf=tuple(sin_leaf),
        angle_ancestors=angle_anc,
        leaf_ancestors=leaf_anc,
        angle_leaves=tuple(tuple(b) for b in below),
        plan_ref=C.byref(plan),
        from_rows=_native.lib().us_from_cartesian_rows,
        to_rows=_native.lib().us_to_cartesian_rows,
        angle_ptrs=(C.c_void_p * n)(),
        ones_leaves=(C.c_uint32 * (n + 1))(*([1] * (n + 1))),
        from_order=tuple((c.s_nodes[s], 1 + s) for s in range(n - 1, -1, -1)),
    )


def plan_for(c: SphericalCoordinates[Any, Any]) -> CompiledTree:
    """Compiled tree of ``c``, built on first use and cached on the object."""
    cached = c._plan
    if cached is None:
        cached = compile_tree(c)
        c._plan = cached
    return cached
