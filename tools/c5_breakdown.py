"""Where the time of BASELINE config 5 goes (create_standard(4), n=96, f = exp(k.x), float64)."""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ultrasphere_b200 as us

dev = torch.device("cuda:0")
n = int(os.environ.get("US_N", "96"))
c = us.create_standard(4)
k = torch.from_numpy(np.random.default_rng(0).uniform(0, 1, 5)).to(dev)
xs, ws = us.roots(c, n, expand_dims_x=True, xp=torch, device=dev, dtype=torch.float64)

def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(reps):
        e0.record(); fn(); e1.record(); e1.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best

nodes = n**3 * 2 * n
res = {"nodes": nodes}
X = c.to_cartesian(xs, as_array=True)
res["to_cartesian_broadcast_ms"] = t(lambda: c.to_cartesian(xs, as_array=True))
ph = torch.einsum("v,v...->...", k, X)
res["einsum_ms"] = t(lambda: torch.einsum("v,v...->...", k, X))
val = torch.exp(ph)
res["exp_ms"] = t(lambda: torch.exp(ph))
wl = [ws[node] for node in c.s_nodes]
res["weighted_reduce_ms"] = t(lambda: us.weighted_reduce(val, wl))
res["weighted_reduce_gbs"] = nodes * 8 / res["weighted_reduce_ms"] / 1e6
res["to_cartesian_dot_ms"] = t(lambda: c.to_cartesian_dot(xs, k))
res["to_cartesian_dot_write_gbs"] = nodes * 8 / res["to_cartesian_dot_ms"] / 1e6
del X, ph, val
ffused = lambda s: torch.exp(c.to_cartesian_dot(s, k))
res["integrate_fused_integrand_ms"] = t(lambda: us.integrate(c, ffused, False, n, xp=torch, device=dev, dtype=torch.float64))
f = lambda s: torch.exp(torch.einsum("v,v...->...", k, c.to_cartesian(s, as_array=True)))
for slab in (0, 16 << 20, 48 << 20, 256 << 20):
    res[f"integrate_slab_{slab>>20}MB_ms"] = t(lambda: us.integrate(c, f, False, n, xp=torch, device=dev, dtype=torch.float64, slab_bytes=slab))
# complex / vector-valued reduction kernels
v3 = torch.randn((n, n, n, 2 * n, 3), device=dev, dtype=torch.float64) if n <= 64 else torch.randn((n, n, n // 2, n, 3), device=dev, dtype=torch.float64)
w3 = [torch.rand(e, device=dev, dtype=torch.float64) for e in v3.shape[:4]]
ms = t(lambda: us.weighted_reduce(v3, w3)); res["reduce_m3_gbs"] = v3.numel() * 8 / ms / 1e6
vc = torch.randn(v3.shape[:4], device=dev, dtype=torch.complex128)
ms = t(lambda: us.weighted_reduce(vc, w3)); res["reduce_complex_gbs"] = vc.numel() * 16 / ms / 1e6
print(json.dumps(res, indent=1))
