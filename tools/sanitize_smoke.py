"""Small run of every kernel family for compute-sanitizer (memcheck): tiled + strided + broadcast
transforms in both dtypes, ragged tails, stacks, reductions, separable, sampler."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ultrasphere_b200 as us
from tests.helpers import make

dev = "cuda:0"
for spec in ("standard:9", "hopf:3", "random:32:0", "spherical"):
    c = make(us, spec)
    for dt in (torch.float32, torch.float64):
        for n in (1000, 1_212_416 + 772, 700_001):  # strided only; tile / ring kernels with a ragged last tile; unaligned (strided throughout)
            x = torch.rand((c.c_ndim, n), device=dev, dtype=dt) * 2 - 1
            s = c.from_cartesian(x)
            y = c.to_cartesian(s, as_array=True)
            assert (y - x).abs().max().item() < 1e-5
c = us.create_standard(4)
k = torch.rand(5, device=dev, dtype=torch.float64)
f = lambda s: torch.exp(torch.einsum("v,v...->...", k, c.to_cartesian(s, as_array=True)))
print(us.integrate(c, f, False, 12, xp=torch, device=dev, dtype=torch.float64).item())
print(us.integrate(c, lambda s: torch.stack([f(s), f(s) * 1j, f(s) + 1], -1), False, 8, xp=torch, device=dev, dtype=torch.float64))
print(us.integrate(c, lambda s: {n: torch.cos(s[n]) ** 2 for n in c.s_nodes}, True, 9, xp=torch, device=dev, dtype=torch.float64))
v = torch.rand((5, 6, 7, 40), device=dev, dtype=torch.float64)
print(us.weighted_reduce(v, [torch.rand(e, device=dev, dtype=torch.float64) for e in (5, 6, 7)]).sum().item())
print(us.random_ball(us.create_standard(20), shape=(5000,), xp=torch, device=dev, dtype=torch.float32).abs().max().item())
print(us.random_ball(c, shape=(5000,), xp=torch, device=dev, type="spherical")["r"].max().item())
torch.cuda.synchronize()
print("sanitize smoke ok")
