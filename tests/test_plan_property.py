"""Property tests of the tree compiler over random branching-type expressions (hypothesis): for any
valid expression the product's factory, node orders and compiled program agree with the oracle, and
the program — walked the way the kernels walk it — reproduces the oracle's numbers bit for bit."""
import numpy as np
from hypothesis import HealthCheck, given, settings, strategies as st

import ultrasphere_b200 as us
from oracle import vks_oracle as vo
from tests.helpers import emulate_from_cartesian, emulate_to_cartesian
from ultrasphere_b200._plan import plan_for


def expressions(max_leaves: int = 24):
    """Valid expressions = pre-order listings of full binary trees: a | b T | b' T | c T T."""
    leaf = st.just("a")

    def extend(children):
        return st.one_of(
            children.map(lambda t: "b" + t),
            children.map(lambda t: "b'" + t),
            st.tuples(children, children).map(lambda p: "c" + p[0] + p[1]),
        )

    return st.recursive(leaf, extend, max_leaves=max_leaves)


@settings(max_examples=80, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(expr=expressions(), seed=st.integers(0, 2**31 - 1))
def test_random_expression_compiles_to_the_oracle(expr, seed):
    c = us.create_from_branching_types(expr)
    t = vo.tree_from_branching(expr)
    assert c.branching_types_expression_str == expr == t.expression_str
    assert list(c.s_nodes) == t.s_nodes and list(c.c_nodes) == t.c_nodes
    assert [list(e) for e in c.cos_edges] == [list(e) for e in t.cos_edges]
    assert {k: v.value for k, v in c.branching_types.items()} == t.branching_types
    ct = plan_for(c)
    # stack depth = the deepest nesting of pending `c` branches
    depth = peak = 0
    toks = vo.parse_branching(expr)
    for i, tok in enumerate(toks):
        if tok == "c":
            depth += 1
            peak = max(peak, depth)
        elif tok == "a" and i != len(toks) - 1:
            depth -= 1
    assert ct.max_stack == peak
    rng = np.random.default_rng(seed)
    x = rng.uniform(-1, 1, (c.c_ndim, 7))
    want = vo.from_cartesian(t, x)
    r, ang = emulate_from_cartesian(ct, list(x))
    assert np.array_equal(r, want["r"])
    for i, node in enumerate(c.s_nodes):
        assert np.array_equal(ang[i], want[node])
    cart = emulate_to_cartesian(ct, ang, r)
    wantc = vo.to_cartesian(t, dict(want))
    for i, leaf in enumerate(c.c_nodes):
        assert np.array_equal(cart[i], wantc[leaf])


@settings(max_examples=40, deadline=None)
@given(n=st.integers(1, 40), seed=st.integers(0, 10_000))
def test_random_factory_trees_compile(n, seed):
    c = us.create_random(n, rng=np.random.default_rng(seed))
    t = vo.create_random(n, rng=np.random.default_rng(seed))
    assert list(c.s_nodes) == t.s_nodes and list(c.c_nodes) == t.c_nodes
    ct = plan_for(c)
    assert ct.n_nodes == n and sorted(ct.angle_slot) == list(range(n))
    leaves = [l for l in ct.cos_leaf + ct.sin_leaf if l >= 0]
    assert sorted(leaves) == list(range(n + 1))
