"""Run one transform config a few times (for ncu): python tools/profile_one.py <spec> <f32|f64> <n>"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ultrasphere_b200 as us
from tests.helpers import make

spec, dt, n = sys.argv[1], sys.argv[2], int(sys.argv[3])
tdt = torch.float32 if dt == "f32" else torch.float64
c = make(us, spec)
x = torch.rand((c.c_ndim, n), device="cuda", dtype=tdt) * 2 - 1
for _ in range(4):
    s = c.from_cartesian(x)
    y = c.to_cartesian(s, as_array=True)
torch.cuda.synchronize()
print("ok", float((y - x).abs().max()))
