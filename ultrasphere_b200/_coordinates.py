"""Coordinate-tree object: the host-side mirror of ultrasphere's ``SphericalCoordinates``.

Public surface (names, argument meaning, dict orders, exceptions) follows the reference so the
class is a drop-in for the hot path:

* tree validation, ``S``, branching types, node orders  — /root/reference/src/ultrasphere/_coordinates.py:29-242, 440-469
* ``from_cartesian`` / ``to_cartesian``                   — _coordinates.py:529-579, 603-656
* ``jacobian``, ``surface_area``, ``volume``              — _coordinates.py:245-308, 668-702

What differs is where the array work happens: the tree is compiled once into a pre-order
program (``_plan.py``) and every transform is ONE launch of a hand-written sm_100a kernel through
the C ABI (``include/ultrasphere_b200.h``).  Inputs must be CUDA ``torch.Tensor``s; there is no
NumPy / CPU path.
"""
from __future__ import annotations

import math
from collections.abc import Iterator, Mapping, Sequence
from enum import StrEnum
from typing import Any, Literal

import networkx as nx

COSSIN = Literal["cos", "sin"]


class BranchingType(StrEnum):
    """Branching type of an internal node (Vilenkin's method of trees)."""

    A = "a"
    B = "b"
    BP = "b'"
    C = "c"


def _successors_by_kind(G: nx.DiGraph, node: Any) -> dict[str, list[Any]]:
    found: dict[str, list[Any]] = {"cos": [], "sin": []}
    for child in G.successors(node):
        kind = G.edges[node, child].get("type")
        if kind in found:
            found[kind].append(child)
    return found


def check_tree(graph: nx.DiGraph, /) -> None:
    """Raise ``ValueError`` unless ``graph`` is a rooted full binary tree whose internal nodes
    have exactly one outgoing ``type="cos"`` and one ``type="sin"`` edge."""
    if not nx.is_tree(graph):
        raise ValueError("The graph is not a tree.")
    if not nx.is_arborescence(graph):
        raise ValueError("The graph is not rooted.")
    for node in graph.nodes:
        degree = graph.out_degree(node)
        if degree not in (0, 2):
            raise ValueError(
                f"Each node should have 0 or 2 successors, but {node} has {degree} successors."
            )
        if degree == 2:
            kinds = {kind for _, _, kind in graph.out_edges(nbunch=node, data="type")}
            if kinds != {"cos", "sin"}:
                raise ValueError(
                    f"Node {node} should have 0 or 2 successors with types 'cos' and 'sin', got {kinds}."
                )


def get_non_leaf_descendants(graph: nx.DiGraph, /) -> dict[Any, int]:
    """Number of internal descendants of every node; NaN for leaves.  Children first."""
    counts: dict[Any, Any] = {}
    for node in reversed(list(nx.topological_sort(graph))):
        kids = list(graph.successors(node))
        counts[node] = sum(counts[k] + 1 for k in kids) if kids else -1
    for node in counts:
        if graph.out_degree(node) == 0:
            counts[node] = math.nan
    return counts


def get_child(G: nx.DiGraph, node: Any, type: COSSIN, /) -> Any:  # noqa: A002
    """The unique child of ``node`` reached through an edge of the given type."""
    hits = [c for c in G.successors(node) if G.edges[node, c]["type"] == type]
    if not hits:
        raise ValueError(f"No {type} child for node {node}.")
    if len(hits) > 1:
        raise ValueError(f"Multiple {type} children for node {node}.")
    return hits[0]


def get_parent(G: nx.DiGraph, node: Any, /) -> Any | None:
    """The parent of ``node`` or ``None`` for the root."""
    parents = list(G.predecessors(node))
    if len(parents) > 1:
        raise ValueError("The node has multiple parents.")
    return parents[0] if parents else None


def _classify(cos_is_leaf: bool, sin_is_leaf: bool) -> BranchingType:
    if cos_is_leaf:
        return BranchingType.A if sin_is_leaf else BranchingType.B
    return BranchingType.BP if sin_is_leaf else BranchingType.C


def get_branching_types(G: nx.DiGraph, /) -> dict[Any, BranchingType]:
    """Branching type of every internal node, in graph node order."""
    out: dict[Any, BranchingType] = {}
    for node in G.nodes:
        if G.out_degree(node) == 0:
            continue
        cos_child = get_child(G, node, "cos")
        sin_child = get_child(G, node, "sin")
        out[node] = _classify(G.out_degree(cos_child) == 0, G.out_degree(sin_child) == 0)
    return out


def get_branching_type_from_digraph(G: nx.DiGraph, /) -> Iterator[BranchingType]:
    """Pre-order walk (cos subtree before sin subtree) yielding branching types."""
    pending = [next(iter(nx.topological_sort(G)))]
    while pending:
        node = pending.pop()
        if G.out_degree(node) == 0:
            raise AssertionError()
        cos_child = get_child(G, node, "cos")
        sin_child = get_child(G, node, "sin")
        kind = _classify(G.out_degree(cos_child) == 0, G.out_degree(sin_child) == 0)
        if kind in (BranchingType.B, BranchingType.C):
            pending.append(sin_child)
        if kind in (BranchingType.BP, BranchingType.C):
            pending.append(cos_child)
        yield kind


def surface_area(c_ndim: int, /, r: float = 1) -> float:
    r"""Surface area :math:`2\pi^{d/2} r^{d-1} / \Gamma(d/2)` of the sphere in ``c_ndim`` dimensions."""
    return 2 * math.pi ** (c_ndim / 2) / math.exp(math.lgamma(c_ndim / 2.0)) * r ** (c_ndim - 1)


def volume(c_ndim: int, /, r: float = 1) -> float:
    r"""Volume :math:`\pi^{d/2} r^d / \Gamma(d/2+1)` of the ball in ``c_ndim`` dimensions."""
    return math.pi ** (c_ndim / 2) / math.exp(math.lgamma(c_ndim / 2.0 + 1.0)) * r**c_ndim


_transform_module: Any = None


def _transform() -> Any:
    """The kernel-launching half of the package, imported on first use (it loads torch and the
    shared library) and remembered: an ``import`` statement per call costs about a microsecond,
    a tenth of what a small transform takes end to end."""
    global _transform_module
    if _transform_module is None:
        from . import _transform as module

        _transform_module = module
    return _transform_module


class SphericalCoordinates[TSpherical, TCartesian]:
    """Polyspherical coordinates described by a rooted cos/sin tree.

    >>> import ultrasphere_b200 as us
    >>> c = us.create_standard(3)
    >>> c
    SphericalCoordinates(bba)
    >>> c.s_nodes, c.c_nodes
    (['theta0', 'theta1', 'theta2'], [0, 1, 2, 3])
    """

    G: nx.DiGraph
    S: Mapping[TSpherical, int]
    branching_types: Mapping[TSpherical, BranchingType]
    s_nodes: Sequence[TSpherical]
    c_nodes: Sequence[TCartesian]
    cos_edges: Sequence[tuple[TSpherical, TCartesian | TSpherical]]
    sin_edges: Sequence[tuple[TSpherical, TCartesian | TSpherical]]

    def __init__(self, tree: nx.DiGraph, /) -> None:
        check_tree(tree)
        self.G = tree
        self.S = get_non_leaf_descendants(tree)
        self.branching_types = get_branching_types(tree)
        order = list(nx.topological_sort(tree))
        try:
            self.c_nodes = sorted(n for n in order if tree.out_degree(n) == 0)
        except TypeError as e:  # unorderable leaf labels (reference dies here with TypeError)
            raise ValueError("Leaf labels must be mutually orderable.") from e
        self.s_nodes = [n for n in order if tree.out_degree(n) == 2]
        self.cos_edges = [e for e in tree.edges if tree.edges[e]["type"] == "cos"]
        self.sin_edges = [e for e in tree.edges if tree.edges[e]["type"] == "sin"]
        self._root = order[0]
        self._plan: Any = None  # compiled lazily by _plan.plan_for(self)
        # leaves are exactly 0..c_ndim-1 in order: a (c_ndim, N) tensor can be addressed row-wise
        self.is_c_keys_sorted_range = list(self.c_nodes) == list(range(len(self.c_nodes)))

    # ---- identity ----------------------------------------------------------------------
    def __hash__(self) -> int:
        return hash(self.G)

    def __eq__(self, other: Any) -> bool:
        return isinstance(other, SphericalCoordinates) and self.G == other.G

    def __getstate__(self) -> dict[Any, Any]:
        return {"G": nx.to_dict_of_dicts(self.G)}

    def __setstate__(self, state: dict[Any, Any]) -> None:
        # the reference rebuilds an undirected graph here and therefore cannot unpickle;
        # keep the directed structure so that round-tripping works
        self.__init__(nx.from_dict_of_dicts(state["G"], create_using=nx.DiGraph))  # type: ignore[misc]

    def __repr__(self) -> str:
        return f"SphericalCoordinates({self.branching_types_expression_str})"

    __str__ = __repr__

    # ---- structure ---------------------------------------------------------------------
    @property
    def root(self) -> Any:
        return self._root

    @property
    def root_index(self) -> int:
        return self.s_nodes.index(self.root)

    @property
    def branching_types_expression(self) -> Sequence[BranchingType]:
        return list(get_branching_type_from_digraph(self.G))

    @property
    def branching_types_expression_str(self) -> str:
        return "".join(t.value for t in self.branching_types_expression)

    @property
    def s_ndim(self) -> int:
        return len(self.s_nodes)

    @property
    def c_ndim(self) -> int:
        return len(self.c_nodes)

    @property
    def is_c_keys_range(self) -> bool:
        return set(self.c_nodes) == set(range(self.c_ndim))

    @property
    def is_s_keys_range(self) -> bool:
        return set(self.s_nodes) == set(range(self.s_ndim))

    def surface_area(self, r: float = 1) -> float:
        return surface_area(self.c_ndim, r=r)

    def volume(self, r: float = 1) -> float:
        return volume(self.c_ndim, r=r)

    # ---- transforms (CUDA) ---------------------------------------------------------------
    def from_cartesian(self, cartesian: Mapping[TCartesian, Any] | Any, /) -> dict[Any, Any]:
        """Cartesian → spherical.  ``cartesian`` is a mapping leaf → tensor, or a tensor whose
        first axis is indexed by leaf id.  Returns ``{"r": ..., <angles, children first>}``.

        One fused kernel launch (``us_from_cartesian``); see ``_transform.from_cartesian``.
        """
        return _transform().from_cartesian(self, cartesian)

    def to_cartesian(self, spherical: Mapping[Any, Any], /, *, as_array: bool = False) -> Any:
        """Spherical → Cartesian.  ``spherical`` maps every angle (and optionally ``"r"``) to a
        tensor; returns a dict in ``c_nodes`` order, or one ``(c_ndim, *shape)`` tensor.

        One fused kernel launch (``us_to_cartesian``); see ``_transform.to_cartesian``.
        """
        return _transform().to_cartesian(self, spherical, as_array=as_array)

    def to_cartesian_dot(self, spherical: Mapping[Any, Any], k: Any, /) -> Any:
        """``sum_v k[v] * x_v`` for ``x = to_cartesian(spherical, as_array=True)`` — the reference
        spells it ``einsum("v,v...->...", k, c.to_cartesian(s, as_array=True))``
        (tests/test_integral.py:88-92) — fused into the transform: the Cartesian grid is never
        written, which matters inside ``integrate`` where it is ``c_ndim`` times the size of the
        integrand.  ``k``: ``(c_ndim,)`` or ``(m, c_ndim)`` in ``c_nodes`` order.  An extension:
        the reference has no such method."""
        from ._transform import to_cartesian_dot

        return to_cartesian_dot(self, spherical, k)

    def jacobian(self, spherical: Mapping[Any, Any], /) -> Any:
        """``r * prod cos^(S_cos+1) sin^(S_sin+1)`` with exponent 0 for leaf children.

        The reference's version (_coordinates.py:668-702) returns NaN for every tree because the
        leaf entries of ``S`` are NaN; this is the formula it intends (a deliberate fix), with the
        reference's own convention of ``r`` to the first power.  One fused kernel launch
        (``us_jacobian``); see ``_transform.jacobian``.
        """
        from ._transform import jacobian

        return jacobian(self, spherical)
