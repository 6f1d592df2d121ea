"""C5 fused integral replayed from a CUDA graph (us.IntegralGraph) with different slab sizes: do L2-sized
slabs (dot -> exp -> reduce stay in the 126 MB L2) beat one pass over the 1.36 GB grid once launches are free?"""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ultrasphere_b200 as us

dev = torch.device("cuda:0")
c4 = us.create_standard(4)
kd = torch.from_numpy(np.random.default_rng(0).uniform(0, 1, 5)).to(dev)
f = lambda s: torch.exp(c4.to_cartesian_dot(s, kd))  # noqa: E731
out = {}
for label, slab in (("one_pass", 0), ("512MB", 512 << 20), ("170MB", 170 << 20), ("85MB", 85 << 20), ("57MB", 57 << 20), ("28MB", 28 << 20), ("14MB", 14 << 20)):
    g = us.IntegralGraph(c4, f, False, 96, xp=torch, device=dev, dtype=torch.float64, slab_bytes=slab)
    for _ in range(3):
        v = g()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        v = g()
    e1.record(); torch.cuda.synchronize()
    out[label] = {"ms": e0.elapsed_time(e1) / 20, "value": v.item()}
    del g
    torch.cuda.empty_cache()
print(json.dumps(out, indent=1))
