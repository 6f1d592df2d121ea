import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ultrasphere_b200 as us
c = us.create_standard(4)
xs, ws = us.roots(c, 96, expand_dims_x=True, xp=torch, device="cuda", dtype=torch.float64)
for _ in range(4):
    X = c.to_cartesian(xs, as_array=True)
torch.cuda.synchronize()
print("ok", X.shape)
