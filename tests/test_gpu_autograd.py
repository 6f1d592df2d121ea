"""torch.autograd through the CUDA transforms (us_to_cartesian_backward / us_from_cartesian_backward):
the reference is differentiable because it is made of array-API calls; here the gradients come from
two dedicated kernels and are checked against numerical Jacobians and against a plain-torch
restatement of the reference's loops (_coordinates.py:555-579, 636-656)."""
import pytest
import torch

import ultrasphere_b200 as us
from tests.helpers import make

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
SPECS = ["polar", "spherical", "standard:3", "standard_prime:3", "hopf:2", "hopf:3", "expr:cbaa", "expr:cbabba", "random:6:3", "random:12:1"]


def leaf_paths(c):
    """For every leaf: [(angle position in s_nodes, True if reached through the sin edge), ...]."""
    paths = {}
    for leaf in c.c_nodes:
        path, node = [], leaf
        while node != c.root:
            parent = us.get_parent(c.G, node)
            path.append((c.s_nodes.index(parent), us.get_child(c.G, parent, "sin") == node))
            node = parent
        paths[leaf] = path[::-1]
    return paths


@pytest.mark.parametrize("spec", SPECS)
def test_to_cartesian_gradients(spec):
    c = make(us, spec)
    gen = torch.Generator(device=DEV).manual_seed(11)
    n = 6
    ang = {k: (torch.rand(n, device=DEV, dtype=torch.float64, generator=gen) * 2.4 + 0.3).requires_grad_() for k in c.s_nodes}
    r = (torch.rand(n, device=DEV, dtype=torch.float64, generator=gen) + 0.5).requires_grad_()
    keys = list(c.s_nodes)

    def f_array(rv, *angles):
        return c.to_cartesian({**dict(zip(keys, angles)), "r": rv}, as_array=True)

    assert torch.autograd.gradcheck(f_array, (r, *[ang[k] for k in keys]), eps=1e-6, atol=1e-8, rtol=1e-6)
    # against the torch restatement, dict output, random cotangents, no r
    paths = leaf_paths(c)
    got = c.to_cartesian(ang)
    w = {leaf: torch.rand(n, device=DEV, dtype=torch.float64, generator=gen) for leaf in c.c_nodes}
    loss = sum((got[leaf] * w[leaf]).sum() for leaf in c.c_nodes)
    grads = torch.autograd.grad(loss, [ang[k] for k in keys])
    ref = 0
    for leaf in c.c_nodes:
        x = torch.ones(n, device=DEV, dtype=torch.float64)
        for s, through_sin in paths[leaf]:
            x = x * (torch.sin(ang[keys[s]]) if through_sin else torch.cos(ang[keys[s]]))
        ref = ref + (x * w[leaf]).sum()
    want = torch.autograd.grad(ref, [ang[k] for k in keys])
    for g, wv, k in zip(grads, want, keys):
        assert torch.allclose(g, wv, rtol=1e-12, atol=1e-13), k


@pytest.mark.parametrize("spec", SPECS)
def test_from_cartesian_gradients(spec):
    c = make(us, spec)
    gen = torch.Generator(device=DEV).manual_seed(13)
    n = 5
    x = (torch.rand((c.c_ndim, n), device=DEV, dtype=torch.float64, generator=gen) * 2 - 1)
    x = torch.where(x.abs() < 0.15, 0.15 * torch.sign(x) + (x == 0) * 0.15, x).requires_grad_()  # away from the atan2 cuts

    def f(xx):
        s = c.from_cartesian(xx)
        return torch.stack([s["r"]] + [s[k] for k in c.s_nodes])

    assert torch.autograd.gradcheck(f, (x,), eps=1e-6, atol=1e-7, rtol=1e-6)
    # mapping input, only some outputs used
    leaves = {k: x[i].detach().clone().requires_grad_() for i, k in enumerate(c.c_nodes)}
    s = c.from_cartesian(leaves)
    first = c.s_nodes[0]
    (s[first].sum() + 2 * s["r"].sum()).backward()
    xd = x.detach().clone().requires_grad_()
    s2 = c.from_cartesian(xd)
    (s2[first].sum() + 2 * s2["r"].sum()).backward()
    for i, k in enumerate(c.c_nodes):
        assert torch.allclose(leaves[k].grad, xd.grad[i], rtol=0, atol=1e-14)


def test_round_trip_gradient_is_the_identity_and_broadcast_inputs_sum():
    c = us.create_hopf(3)
    gen = torch.Generator(device=DEV).manual_seed(3)
    x = (torch.rand((8, 4, 7), device=DEV, dtype=torch.float64, generator=gen) + 0.2).requires_grad_()
    w = torch.rand((8, 4, 7), device=DEV, dtype=torch.float64, generator=gen)
    y = c.to_cartesian(c.from_cartesian(x), as_array=True)
    (y * w).sum().backward()
    assert torch.allclose(x.grad, w, rtol=1e-11, atol=1e-12)
    # broadcast operands: gradients are summed back to each operand's own shape
    s3 = us.create_standard(3)
    th = {"theta0": torch.rand((5, 1, 1), device=DEV, dtype=torch.float64, generator=gen).requires_grad_(),
          "theta1": torch.rand((1, 6, 1), device=DEV, dtype=torch.float64, generator=gen).requires_grad_(),
          "theta2": torch.rand((1, 1, 7), device=DEV, dtype=torch.float64, generator=gen).requires_grad_(),
          "r": torch.tensor(2.0, device=DEV, dtype=torch.float64, requires_grad=True)}
    out = s3.to_cartesian(th, as_array=True)
    assert out.shape == (4, 5, 6, 7)
    out.sum().backward()
    assert th["theta0"].grad.shape == (5, 1, 1) and th["theta2"].grad.shape == (1, 1, 7) and th["r"].grad.shape == ()
    eager = torch.stack([
        th["r"] * torch.cos(th["theta0"]) * torch.ones(1, 6, 7, device=DEV, dtype=torch.float64),
        th["r"] * torch.sin(th["theta0"]) * torch.cos(th["theta1"]) * torch.ones(1, 1, 7, device=DEV, dtype=torch.float64),
        th["r"] * torch.sin(th["theta0"]) * torch.sin(th["theta1"]) * torch.cos(th["theta2"]),
        th["r"] * torch.sin(th["theta0"]) * torch.sin(th["theta1"]) * torch.sin(th["theta2"]),
    ])
    want = torch.autograd.grad(eager.sum(), [th["theta0"], th["theta1"], th["theta2"], th["r"]])
    for g, wv in zip([th["theta0"].grad, th["theta1"].grad, th["theta2"].grad, th["r"].grad], want):
        assert torch.allclose(g, wv, rtol=1e-12, atol=1e-12)


def test_float32_gradients_dot_and_jacobian():
    c = us.create_spherical()
    x = torch.rand((3, 1000), device=DEV, requires_grad=True)
    s = c.from_cartesian(x)
    assert s["r"].requires_grad and s["r"].dtype == torch.float32
    y = c.to_cartesian(s, as_array=True)
    y.pow(2).sum().backward()
    assert torch.allclose(x.grad, 2 * x.detach(), rtol=2e-4, atol=2e-5)
    # to_cartesian_dot under autograd = the contraction of the differentiable transform
    th = {k: v.detach().double().requires_grad_() for k, v in s.items()}
    k = torch.tensor([0.3, -1.2, 0.5], device=DEV, dtype=torch.float64, requires_grad=True)
    d = c.to_cartesian_dot(th, k)
    d.sum().backward()
    xs = c.to_cartesian({kk: v.detach() for kk, v in th.items()}, as_array=True)
    assert torch.allclose(k.grad, xs.sum(dim=1), rtol=1e-12)
    assert all(v.grad is not None for v in th.values())
    jac = c.jacobian(th)  # under autograd: the same product in torch ops
    with torch.no_grad():
        fused = c.jacobian(th)
    assert jac.requires_grad and torch.allclose(jac, fused, rtol=1e-13)
    with torch.no_grad():
        assert not c.from_cartesian(x)["r"].requires_grad
