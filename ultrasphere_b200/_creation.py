"""Factories for coordinate trees — same names, arguments and node labels as the reference
(/root/reference/src/ultrasphere/_creation.py:12-405).

A branching-type expression is the pre-order listing (cos subtree first) of the internal nodes'
types; it is turned into a ``networkx.DiGraph`` here only because ``SphericalCoordinates.G`` is
part of the public surface.  The device never sees the graph: ``_plan.py`` flattens the same
expression into the program the kernels interpret.
"""
from __future__ import annotations

from collections.abc import Sequence
from typing import Any

import networkx as nx
import numpy as np

from ._coordinates import BranchingType, SphericalCoordinates


def _tokenize(expr: str | Sequence[BranchingType | str]) -> list[BranchingType]:
    if not isinstance(expr, str):
        return [BranchingType(t) for t in expr]
    text = expr.replace("bp", "b'")
    tokens: list[BranchingType] = []
    pos = 0
    while pos < len(text):
        if text.startswith("b'", pos):
            tokens.append(BranchingType.BP)
            pos += 2
        elif text[pos] in "abc":
            tokens.append(BranchingType(text[pos]))
            pos += 1
        else:
            raise ValueError(f"Invalid branching type: {text[pos:]}")
    return tokens


def _s_node_name_default(idx: int) -> Any:
    return f"theta{idx}"


def _e_node_name_default(idx: int) -> Any:
    return idx


def _get_digraph_from_branching_type(branching_types: str | Sequence[BranchingType | str]) -> nx.DiGraph:
    """Angles are named ``theta{k}`` and leaves numbered ``0,1,…`` in creation order; a ``c`` node
    parks its sin child until the branch being written is closed by an ``a``."""
    tokens = _tokenize(branching_types)
    G = nx.DiGraph()
    names = (_s_node_name_default(i) for i in range(len(tokens) + 2))
    leaf_ids = iter(range(len(tokens) + 2))
    parked: list[Any] = []
    node = next(names)
    G.add_node(node)

    def leaf(parent: Any, kind: str) -> None:
        child = _e_node_name_default(next(leaf_ids))
        G.add_node(child)
        G.add_edge(parent, child, type=kind)

    def angle(parent: Any, kind: str) -> Any:
        child = next(names)
        G.add_node(child)
        G.add_edge(parent, child, type=kind)
        return child

    for pos, tok in enumerate(tokens):
        if tok is BranchingType.A:
            leaf(node, "cos")
            leaf(node, "sin")
            if pos == len(tokens) - 1:
                break
            if not parked:
                raise ValueError("Invalid branching types.")
            node = parked.pop()
        elif tok is BranchingType.B:
            leaf(node, "cos")
            node = angle(node, "sin")
        elif tok is BranchingType.BP:
            nxt = angle(node, "cos")
            leaf(node, "sin")
            node = nxt
        else:
            nxt = angle(node, "cos")
            parked.append(angle(node, "sin"))
            node = nxt
    return G


def create_from_branching_types(
    branching_types: str | Sequence[BranchingType | str],
) -> SphericalCoordinates[Any, Any]:
    """Coordinates from an expression such as ``"ba"`` (``"bp"`` is accepted for ``"b'"``)."""
    return SphericalCoordinates(_get_digraph_from_branching_type(branching_types))


def create_polar() -> SphericalCoordinates[Any, Any]:
    """x0 = r cos φ, x1 = r sin φ."""
    G = _get_digraph_from_branching_type("a")
    return SphericalCoordinates(nx.relabel_nodes(G, {"theta0": "phi"}))


def create_spherical() -> SphericalCoordinates[Any, Any]:
    """x0 = r sin θ cos φ, x1 = r sin θ sin φ, x2 = r cos θ."""
    G = _get_digraph_from_branching_type("ba")
    relabel = {0: 2, 2: 1, 1: 0, "theta0": "theta", "theta1": "phi"}
    return SphericalCoordinates(nx.relabel_nodes(G, relabel))


def create_standard(s_ndim: int) -> SphericalCoordinates[Any, Any]:
    """x0 = cos θ0, x1 = sin θ0 cos θ1, …  (expression ``b…ba``)."""
    return create_from_branching_types("b" * (s_ndim - 1) + "a" if s_ndim > 0 else "")


def create_standard_prime(s_ndim: int) -> SphericalCoordinates[Any, Any]:
    """x0 = sin θ0, x1 = cos θ0 sin θ1, …  (expression ``b'…b'a``)."""
    return create_from_branching_types("bp" * (s_ndim - 1) + "a" if s_ndim > 0 else "")


def _hopf_expression(q: int) -> str:
    if q < 0:
        raise ValueError("q should be non-negative.")
    expr = "a" if q >= 1 else ""
    for _ in range(q - 1):
        expr = "c" + expr + expr
    return expr


def create_hopf(q: int) -> SphericalCoordinates[Any, Any]:
    """Hopf coordinates with ``2**q`` Cartesian dimensions (perfect binary tree)."""
    return create_from_branching_types(_hopf_expression(q))


def _get_random_digraph(s_ndim: int, *, rng: np.random.Generator | None = None) -> nx.DiGraph:
    """Split a uniformly chosen leaf ``s_ndim`` times, then relabel leaves ``0..`` in leaf-list
    order and internal nodes ``theta{i}`` in set-iteration order of their ids — the labelling the
    reference produces for the same ``rng``."""
    rng = np.random.default_rng() if rng is None else rng
    G = nx.DiGraph()
    G.add_node(0)
    leaves = [0]
    for _ in range(s_ndim):
        parent = rng.choice(leaves)
        leaves.remove(parent)
        for kind in ("cos", "sin"):
            child = len(G)
            G.add_node(child)
            G.add_edge(parent, child, type=kind)
            leaves.append(child)
    internal = set(G.nodes) - set(leaves)
    mapping = {n: _e_node_name_default(i) for i, n in enumerate(leaves)}
    mapping |= {n: _s_node_name_default(i) for i, n in enumerate(internal)}
    return nx.relabel_nodes(G, mapping)


def create_random(s_ndim: int, *, rng: np.random.Generator | None = None) -> SphericalCoordinates[Any, Any]:
    """A random tree with ``s_ndim`` angles."""
    return SphericalCoordinates(_get_random_digraph(s_ndim, rng=rng))
