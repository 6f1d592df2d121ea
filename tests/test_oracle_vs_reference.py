"""Live comparison of the NumPy oracle with the unmodified reference.  Runs only where
/root/reference is mounted (the build container); the GPU box relies on tests/golden."""
import numpy as np
import pytest

from oracle import refload, vks_oracle as vo
from tests.helpers import make

pytestmark = pytest.mark.skipif(not refload.available(), reason="/root/reference not mounted")

SPECS = [
    "polar", "spherical", "standard:1", "standard:4", "standard:9", "standard_prime:8", "hopf:2", "hopf:3",
    "random:5:0", "random:32:0", "random:40:11", "expr:cbaa", "expr:cbabba",
]


@pytest.fixture(scope="module")
def ref():
    return refload.load().us


@pytest.mark.parametrize("spec", SPECS)
def test_bit_exact_transforms(ref, spec):
    c, t = make(ref, spec), vo.make(spec)
    assert list(c.s_nodes) == t.s_nodes and list(c.c_nodes) == t.c_nodes
    rng = np.random.default_rng(1)
    for dt in (np.float64, np.float32):
        x = rng.uniform(-1, 1, (c.c_ndim, 4, 65)).astype(dt)
        a, b = c.from_cartesian(x), vo.from_cartesian(t, x)
        assert list(a) == list(b)
        for k in a:
            assert np.array_equal(a[k], b[k])
        ca, cb = c.to_cartesian(a), vo.to_cartesian(t, b)
        assert list(ca) == list(cb)
        for k in ca:
            assert np.array_equal(ca[k], cb[k])


@pytest.mark.parametrize("spec", ["spherical", "standard:3", "hopf:2", "random:5:0", "expr:cbaa"])
def test_bit_exact_quadrature(ref, spec):
    c, t = make(ref, spec), vo.make(spec)
    k = np.random.default_rng(2).uniform(0, 1, c.c_ndim)
    for n in (3, 6, 9):
        xa, wa = ref.roots(c, n, expand_dims_x=True, expand_dims_w=True, xp=np)
        xb, wb = vo.roots(t, n, expand_dims_x=True, expand_dims_w=True)
        for node in c.s_nodes:
            assert np.array_equal(xa[node], xb[node]) and np.array_equal(wa[node], wb[node])
        fa = lambda s: np.exp(np.einsum("v,v...->...", k, c.to_cartesian(s, as_array=True)))  # noqa: E731
        fb = lambda s: np.exp(np.einsum("v,v...->...", k, vo.to_cartesian(t, s, as_array=True)))  # noqa: E731
        assert ref.integrate(c, fa, False, n, xp=np) == vo.integrate(t, fb, False, n)
