"""C5 pieces for ncu: the broadcast grid kernel and the TMA-streamed reduction (n=96, float64)."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ultrasphere_b200 as us
c = us.create_standard(4)
xs, ws = us.roots(c, 96, expand_dims_x=True, xp=torch, device="cuda", dtype=torch.float64)
wl = [w.ravel() for w in ws.values()]
k = torch.rand(5, device="cuda", dtype=torch.float64)
for _ in range(3):
    X = c.to_cartesian(xs, as_array=True)
    D = c.to_cartesian_dot(xs, k)
    val = torch.exp(X[0])
    out = us.weighted_reduce(val, wl)
torch.cuda.synchronize()
print("ok", out.item())
