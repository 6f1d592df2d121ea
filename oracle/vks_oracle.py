"""CPU oracle for the ultrasphere hot path — a NumPy restatement.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product package
(``ultrasphere_b200``) never does: its transforms and reductions run in the CUDA library or
fail loudly.

What is restated (all citations relative to /root/reference/src/ultrasphere):

* tree construction from a branching-type string             _creation.py:12-97
* the named factories and the random tree                    _creation.py:100-405
* node orderings, S, branching types, edge lists             _coordinates.py:81-109,172-242,440-469
* ``from_cartesian`` / ``to_cartesian``                        _coordinates.py:529-579, 603-656
* ``roots`` / ``integrate``                                    _integral.py:18-116, 157-276

The tree is held in plain dicts and lists (no networkx); the one place the reference depends
on networkx semantics — ``nx.topological_sort`` — is restated as the level-by-level order that
``nx.topological_generations`` (networkx 3.6.1, algorithms/dag.py) produces on a rooted tree:
generation k = nodes at depth k, inside a generation in the order their parents were scanned,
children in adjacency-insertion order.

Arithmetic that lives in third-party code is *called*, not re-derived, exactly as the
reference calls it: ``scipy.special.roots_jacobi`` (scipy 1.18.x) for the 1-D Gauss–Jacobi
rules, and NumPy's ``sin/cos/arctan2/linalg.vector_norm/vecdot`` ufuncs.

Parity pin: ``tests/golden/*.npz|json`` were produced by ``oracle/gen_golden.py`` from the
unmodified reference (``oracle/refload.py``) in the build container; ``tests/test_oracle_golden.py``
checks this module against them bit-for-bit, and ``tests/test_oracle_vs_reference.py`` re-runs the
comparison live whenever /root/reference is mounted.
"""
from __future__ import annotations

import math
from collections.abc import Callable, Mapping, Sequence
from typing import Any

import numpy as np

A, B, BP, C = "a", "b", "b'", "c"


# --------------------------------------------------------------------------------------
# tree
# --------------------------------------------------------------------------------------
class OracleTree:
    """A rooted full binary tree with one "cos" and one "sin" edge per internal node.

    ``adj[node]`` is the list of ``(child, kind)`` in insertion order (empty for leaves);
    ``order`` is the node insertion order (what iterating a DiGraph yields).
    """

    def __init__(self, order: Sequence[Any], adj: Mapping[Any, Sequence[tuple[Any, str]]]):
        self.order = list(order)
        self.adj = {k: list(adj.get(k, ())) for k in self.order}
        indeg = {k: 0 for k in self.order}
        for k in self.order:
            for ch, _ in self.adj[k]:
                indeg[ch] += 1
        # _coordinates.py:52-78 (check_tree): one root, 0 or 2 successors, {"cos","sin"}
        roots = [k for k in self.order if indeg[k] == 0]
        if len(roots) != 1 or any(v > 1 for v in indeg.values()):
            raise ValueError("The graph is not a rooted tree.")
        for k in self.order:
            kinds = sorted(kind for _, kind in self.adj[k])
            if kinds not in ([], ["cos", "sin"]):
                raise ValueError(f"Node {k} should have 0 or 2 successors of types cos/sin.")
        # nx.topological_sort == concatenated generations (see module docstring)
        topo: list[Any] = []
        gen = roots
        while gen:
            topo.extend(gen)
            gen = [ch for node in gen for ch, _ in self.adj[node]]
        if len(topo) != len(self.order):
            raise ValueError("The graph is not a tree.")
        self.topo = topo
        self.root = topo[0]
        # _coordinates.py:466-467
        self.c_nodes = sorted(k for k in topo if not self.adj[k])
        self.s_nodes = [k for k in topo if self.adj[k]]
        # _coordinates.py:468-469: G.edges iterates node order, then adjacency order
        self.cos_edges = [(k, ch) for k in self.order for ch, kind in self.adj[k] if kind == "cos"]
        self.sin_edges = [(k, ch) for k in self.order for ch, kind in self.adj[k] if kind == "sin"]
        # _coordinates.py:96-109 (S: number of non-leaf descendants; NaN on leaves)
        cnt: dict[Any, int] = {}
        for k in reversed(topo):
            cnt[k] = -1 if not self.adj[k] else sum(cnt[ch] + 1 for ch, _ in self.adj[k])
        self.S = {k: (cnt[k] if self.adj[k] else float("nan")) for k in cnt}
        # _coordinates.py:187-203
        self.branching_types = {k: self._type(k) for k in self.order if self.adj[k]}

    def child(self, node: Any, kind: str) -> Any:
        """_coordinates.py:112-139"""
        hits = [ch for ch, t in self.adj[node] if t == kind]
        if len(hits) != 1:
            raise ValueError(f"No unique {kind} child for node {node}.")
        return hits[0]

    def _type(self, node: Any) -> str:
        cos_leaf = not self.adj[self.child(node, "cos")]
        sin_leaf = not self.adj[self.child(node, "sin")]
        if cos_leaf and sin_leaf:
            return A
        if cos_leaf:
            return B
        if sin_leaf:
            return BP
        return C

    @property
    def s_ndim(self) -> int:
        return len(self.s_nodes)

    @property
    def c_ndim(self) -> int:
        return len(self.c_nodes)

    @property
    def expression(self) -> list[str]:
        """Pre-order walk, cos subtree before sin subtree (_coordinates.py:220-242)."""
        out: list[str] = []
        stack = [self.root]
        while stack:
            node = stack.pop()
            if not self.adj[node]:
                raise AssertionError()
            t = self.branching_types[node]
            if t == B:
                stack.append(self.child(node, "sin"))
            elif t == BP:
                stack.append(self.child(node, "cos"))
            elif t == C:
                stack.extend([self.child(node, "sin"), self.child(node, "cos")])
            out.append(t)
        return out

    @property
    def expression_str(self) -> str:
        return "".join(self.expression)

    def relabel(self, mapping: Mapping[Any, Any]) -> "OracleTree":
        """What nx.relabel_nodes(copy=True) does to node and adjacency order."""
        m = lambda k: mapping.get(k, k)  # noqa: E731
        return OracleTree(
            [m(k) for k in self.order],
            {m(k): [(m(ch), t) for ch, t in v] for k, v in self.adj.items()},
        )


def parse_branching(expr: str | Sequence[str]) -> list[str]:
    """_creation.py:35-48: "bp" is an alias of "b'"; anything else raises ValueError."""
    if not isinstance(expr, str):
        return [str(getattr(t, "value", t)) for t in expr]
    s = expr.replace("bp", "b'")
    out = []
    while s:
        if s.startswith("b'"):
            out.append(BP)
            s = s[2:]
        elif s[0] in "abc":
            out.append(s[0])
            s = s[1:]
        else:
            raise ValueError(f"Invalid branching type: {s}")
    return out


def tree_from_branching(expr: str | Sequence[str]) -> OracleTree:
    """_creation.py:50-97: `c` defers its sin child on a stack; `a` closes a branch and pops."""
    types = parse_branching(expr)
    order: list[Any] = []
    adj: dict[Any, list[tuple[Any, str]]] = {}

    def add(node: Any) -> None:
        if node not in adj:
            adj[node] = []
            order.append(node)

    pending: list[Any] = []
    n_leaf = 0
    n_ang = 1
    cur: Any = "theta0"
    add(cur)
    for i, t in enumerate(types):
        if t == A:
            for kind in ("cos", "sin"):
                add(n_leaf)
                adj[cur].append((n_leaf, kind))
                n_leaf += 1
            if i == len(types) - 1:
                break
            if not pending:
                raise ValueError("Invalid branching types.")
            nxt = pending.pop()
        elif t == B:
            add(n_leaf)
            adj[cur].append((n_leaf, "cos"))
            n_leaf += 1
            nxt = f"theta{n_ang}"
            add(nxt)
            adj[cur].append((nxt, "sin"))
            n_ang += 1
        elif t == BP:
            nxt = f"theta{n_ang}"
            add(nxt)
            adj[cur].append((nxt, "cos"))
            n_ang += 1
            add(n_leaf)
            adj[cur].append((n_leaf, "sin"))
            n_leaf += 1
        elif t == C:
            nxt = f"theta{n_ang}"
            add(nxt)
            adj[cur].append((nxt, "cos"))
            n_ang += 1
            deferred = f"theta{n_ang}"
            add(deferred)
            adj[cur].append((deferred, "sin"))
            pending.append(deferred)
            n_ang += 1
        else:
            raise ValueError(f"Invalid branching type: {t}")
        cur = nxt
    return OracleTree(order, adj)


def create_polar() -> OracleTree:
    """_creation.py:124-126"""
    return tree_from_branching("a").relabel({"theta0": "phi"})


def create_spherical() -> OracleTree:
    """_creation.py:156-159: x0 = r sinθ cosφ, x1 = r sinθ sinφ, x2 = r cosθ"""
    return tree_from_branching("ba").relabel({0: 2, 2: 1, 1: 0, "theta0": "theta", "theta1": "phi"})


def create_standard(s_ndim: int) -> OracleTree:
    """_creation.py:194-196"""
    return tree_from_branching("" if s_ndim == 0 else "b" * (s_ndim - 1) + "a")


def create_standard_prime(s_ndim: int) -> OracleTree:
    """_creation.py:231-233"""
    return tree_from_branching("" if s_ndim == 0 else "bp" * (s_ndim - 1) + "a")


def hopf_expression(q: int) -> str:
    """_creation.py:262-270"""
    if q < 0:
        raise ValueError("q should be non-negative.")
    if q == 0:
        return ""
    if q == 1:
        return "a"
    h = hopf_expression(q - 1)
    return "c" + h + h


def create_hopf(q: int) -> OracleTree:
    return tree_from_branching(hopf_expression(q))


def create_random(s_ndim: int, *, rng: np.random.Generator | None = None) -> OracleTree:
    """_creation.py:350-370: split a uniformly chosen leaf s_ndim times, then relabel.

    Leaves are renumbered in ``leaf_nodes`` list order and internal nodes become ``theta{i}``
    in *set iteration order* of their integer ids (kept as a Python set here on purpose).
    """
    rng = np.random.default_rng() if rng is None else rng
    order: list[Any] = [0]
    adj: dict[Any, list[tuple[Any, str]]] = {0: []}
    leaves: list[Any] = [0]
    for _ in range(s_ndim):
        parent = rng.choice(leaves)
        leaves.remove(parent)
        parent = int(parent)
        for kind in ("cos", "sin"):
            node = len(order)
            order.append(node)
            adj[node] = []
            adj[parent].append((node, kind))
            leaves.append(node)
    internal = set(order) - set(leaves)
    mapping = {node: i for i, node in enumerate(leaves)} | {
        node: f"theta{i}" for i, node in enumerate(internal)
    }
    return OracleTree(order, adj).relabel(mapping)


# --------------------------------------------------------------------------------------
# transforms (NumPy, same call sequence as the reference)
# --------------------------------------------------------------------------------------
def from_cartesian(tree: OracleTree, cartesian: Any) -> dict[Any, np.ndarray]:
    """_coordinates.py:555-579.  `cartesian` is a mapping or anything indexable by leaf id."""
    leaves = [cartesian[k] for k in tree.c_nodes]
    if tree.s_ndim > 0:
        r = np.linalg.vector_norm(np.stack(np.broadcast_arrays(*leaves)), axis=0)
    else:
        r = leaves[0]
    result: dict[Any, np.ndarray] = {"r": r}
    tmp = dict(zip(tree.c_nodes, leaves))
    for node in reversed(tree.topo):
        if not tree.adj[node]:
            continue
        cos = tmp[tree.child(node, "cos")]
        sin = tmp[tree.child(node, "sin")]
        result[node] = np.atan2(sin, cos)
        tmp[node] = np.linalg.vector_norm(np.stack(np.broadcast_arrays(sin, cos)), axis=0)
    return result


def to_cartesian(tree: OracleTree, spherical: Mapping[Any, Any], *, as_array: bool = False) -> Any:
    """_coordinates.py:636-656.  Multiply order is root→leaf; r defaults to the int 1."""
    for k in tree.s_nodes:
        spherical[k]  # KeyError for a missing angle, as `array_namespace(*[...])` would raise
    result: dict[Any, Any] = {tree.root: spherical.get("r", 1)}
    for node in tree.topo:
        if not tree.adj[node]:
            continue
        cos = np.cos(spherical[node])
        sin = np.sin(spherical[node])
        result[tree.child(node, "cos")] = result[node] * cos
        result[tree.child(node, "sin")] = result[node] * sin
    result = {k: result[k] for k in tree.c_nodes}
    if as_array:
        return np.stack(np.broadcast_arrays(*[result[k] for k in tree.c_nodes]), axis=0)
    return result


# --------------------------------------------------------------------------------------
# quadrature
# --------------------------------------------------------------------------------------
def roots(
    tree: OracleTree,
    n: int,
    *,
    expand_dims_x: bool,
    expand_dims_w: bool = False,
    dtype: Any | None = None,
) -> tuple[dict[Any, np.ndarray], dict[Any, np.ndarray]]:
    """_integral.py:81-116 with xp = numpy."""
    from scipy.special import roots_jacobi

    xs: dict[Any, np.ndarray] = {}
    ws: dict[Any, np.ndarray] = {}
    for i, node in enumerate(tree.s_nodes):
        t = tree.branching_types[node]
        if t == A:
            x = np.arange(2 * n, dtype=dtype) * np.pi / n
            w = np.ones(2 * n, dtype=dtype) * np.pi / n
        elif t == B:
            beta = tree.S[tree.child(node, "sin")] / 2
            x, w = roots_jacobi(n, beta, beta)
            x = np.acos(x)
        elif t == BP:
            alpha = tree.S[tree.child(node, "cos")] / 2
            x, w = roots_jacobi(n, alpha, alpha)
            x = np.asin(x)
        elif t == C:
            # NB the reference passes (S_cos/2, S_sin/2) as (alpha, beta) — SURVEY.md App. A.1;
            # parity means reproducing that call, not "fixing" it.
            alpha = tree.S[tree.child(node, "cos")] / 2
            beta = tree.S[tree.child(node, "sin")] / 2
            x, w = roots_jacobi(n, alpha, beta)
            w /= 2 ** (alpha + beta + 2)
            x = np.acos(x) / 2
        else:
            raise ValueError(f"Invalid branching type {t}.")
        x = np.asarray(x, dtype=dtype)
        w = np.asarray(w, dtype=dtype)
        idx = (None,) * i + (slice(None),) + (None,) * (tree.s_ndim - i - 1)
        if expand_dims_x:
            x = x[idx]
        if expand_dims_w:
            w = w[idx]
        xs[node] = x
        ws[node] = w
    return xs, ws


def contract(tree: OracleTree, val: Any, ws: Mapping[Any, np.ndarray]) -> Any:
    """The weighted contraction of _integral.py:245-276 on already-evaluated values."""
    if isinstance(val, Mapping):
        result = {}
        for node in tree.s_nodes:
            value = val[node]
            np.broadcast_shapes(value.shape[:1], ws[node].shape)
            w = np.reshape(ws[node], (-1,) + (1,) * (value.ndim - 1))
            if value.shape[0] == 1:
                result[node] = value[0, ...] * np.sum(w)
            else:
                result[node] = np.vecdot(w, value, axis=0)
        return result
    if val.ndim < tree.s_ndim:
        raise ValueError(
            f"The dimension of the return value of f should be at least {tree.s_ndim}, got {val.ndim}."
        )
    for node in tree.s_nodes:
        w = ws[node]
        if val.shape[0] == 1:
            val = val[0, ...] * np.sum(w)
        else:
            val = np.vecdot(w[(slice(None),) + (None,) * (val.ndim - 1)], val, axis=0)
    return val


def integrate(
    tree: OracleTree,
    f: Callable[[Mapping[Any, np.ndarray]], Any] | Any,
    does_f_support_separation_of_variables: bool,
    n: int,
    *,
    dtype: Any | None = None,
) -> Any:
    """_integral.py:228-276 with xp = numpy."""
    xs, ws = roots(tree, n, expand_dims_x=not does_f_support_separation_of_variables, dtype=dtype)
    val = f(xs) if callable(f) else f
    return contract(tree, val, ws)


# --------------------------------------------------------------------------------------
# helpers for tests / bench
# --------------------------------------------------------------------------------------
FACTORIES: dict[str, Callable[..., OracleTree]] = {
    "polar": create_polar,
    "spherical": create_spherical,
    "standard": create_standard,
    "standard_prime": create_standard_prime,
    "hopf": create_hopf,
    "from_branching_types": tree_from_branching,
}


def make(spec: str) -> OracleTree:
    """'standard:9', 'hopf:3', 'random:32:0' (seed), 'spherical', 'expr:cbaa' → tree."""
    head, *args = spec.split(":")
    if head == "random":
        return create_random(int(args[0]), rng=np.random.default_rng(int(args[1])))
    if head == "expr":
        return tree_from_branching(args[0] if args else "")
    if head in ("polar", "spherical"):
        return FACTORIES[head]()
    return FACTORIES[head](int(args[0]))


def ulp_distance(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """Distance in units-in-the-last-place between same-dtype float arrays (monotone int map)."""
    a = np.asarray(a)
    b = np.asarray(b)
    assert a.dtype == b.dtype and a.dtype in (np.float32, np.float64)
    a, b = np.broadcast_arrays(a, b)
    it = np.int32 if a.dtype == np.float32 else np.int64
    bits = 31 if a.dtype == np.float32 else 63
    ia = np.ascontiguousarray(a).view(it).astype(np.int64)
    ib = np.ascontiguousarray(b).view(it).astype(np.int64)
    mask = np.int64((1 << bits) - 1)
    ma, mb = ia & mask, ib & mask  # magnitudes (sign bit stripped), exact in int64
    same = (ia < 0) == (ib < 0)
    # same sign: exact integer difference; opposite signs: ma+mb (float is exact while < 2^53,
    # and merely "huge" beyond that, which is all a tolerance check needs)
    d = np.where(same, np.abs(ma - mb).astype(np.float64), ma.astype(np.float64) + mb.astype(np.float64))
    nan = np.isnan(a) | np.isnan(b)
    both = np.isnan(a) & np.isnan(b)
    return np.where(both, 0.0, np.where(nan, np.inf, d))


def analytic_exp_kx(k: np.ndarray) -> float:
    """∫_{S^{d-1}} exp(k·x) dω = (2π)^{d/2} |k|^{1-d/2} I_{d/2-1}(|k|)   (SURVEY.md §8c)."""
    from scipy.special import iv

    d = len(k)
    kn = float(np.linalg.norm(k))
    return (2 * math.pi) ** (d / 2) * kn ** (1 - d / 2) * float(iv(d / 2 - 1, kn))
