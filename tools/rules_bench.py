"""Latency of building the 1-D quadrature tables (roots()): host scipy + H2D copies (the reference's
way, cold and with the per-(n, alpha, beta) host cache warm) against one launch of
us_quadrature_rules, and what that does to a small integral whose n changes every call.
Prints one JSON object."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ultrasphere_b200 as us
from ultrasphere_b200 import _integral

dev = torch.device("cuda:0")

def wall(fn, reps):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    return best * 1e6

def dev_time(fn, reps=20):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(reps):
        e0.record(); fn(); e1.record(); e1.synchronize(); best = min(best, e0.elapsed_time(e1) * 1e3)
    return best

out = {}
for spec, c in (("standard:4", us.create_standard(4)), ("hopf:3", us.create_hopf(3)), ("standard:9", us.create_standard(9))):
    for n in (8, 32, 96, 512):
        kw = dict(expand_dims_x=True, xp=torch, device=dev, dtype=torch.float64)
        def cold():
            _integral._jacobi_rule.cache_clear(); us.roots(c, n, **kw)
        def warm():
            us.roots(c, n, **kw)
        def device():
            with us.rule_backend("device"):
                us.roots(c, n, **kw)
        row = {"scipy_cold_us": wall(cold, 5), "scipy_cached_us": wall(warm, 20), "device_us": wall(device, 20),
               "device_kernel_us": dev_time(device)}
        out[f"roots|{spec}|n{n}"] = row

# the ABI call alone: 200 back-to-back launches for standard:4 (three Gauss axes + the trapezoid)
import ctypes as C
from ultrasphere_b200 import _native
for n in (8, 96, 512):
    kinds, counts = [1, 1, 1, 0], [n, n, n, 2 * n]
    alphas = [1.0, 0.5, 0.0, 0.0]
    buf = [torch.empty(2 * n, dtype=torch.float64, device=dev) for _ in range(8)]
    args = (4, (C.c_uint8 * 4)(*kinds), (C.c_int64 * 4)(*counts), (C.c_double * 4)(*alphas), (C.c_double * 4)(*alphas),
            (C.c_int32 * 4)(1, 1, 1, 1), (C.c_void_p * 4)(*[b.data_ptr() for b in buf[:4]]), (C.c_void_p * 4)(*[b.data_ptr() for b in buf[4:]]),
            C.c_void_p(torch.cuda.current_stream().cuda_stream))
    lib = _native.lib()
    def burst():
        for _ in range(200):
            lib.us_quadrature_rules(*args)
    out[f"abi|standard:4|n{n}"] = {"wall_us_per_call": wall(burst, 3) / 200, "device_us_per_call": dev_time(burst, 3) / 200}

# many small integrals with changing n: surface measure of exp(k.x) on S^4, n cycling so that no table is reused
c = us.create_standard(4)
k = torch.tensor(np.random.default_rng(0).uniform(0, 1, 5), device=dev)
f = lambda s: torch.exp(torch.einsum("v,v...->...", k, c.to_cartesian(s, as_array=True)))
ns = list(range(6, 30))
def sweep(backend):
    _integral._jacobi_rule.cache_clear(); c.__dict__.pop("_quad_cache", None)
    with us.rule_backend(backend):
        return [us.integrate(c, f, False, n, xp=torch, device=dev, dtype=torch.float64) for n in ns]
for backend in ("scipy", "device"):
    sweep(backend); torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        t0 = time.perf_counter(); vals = sweep(backend); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    out[f"integrate_sweep|standard:4|n6..29|{backend}"] = {"us_per_integral": best * 1e6 / len(ns), "last": vals[-1].item()}
# one small integral (n = 16, fused integrand): eager call against CUDA-graph replay
kk = torch.tensor(np.random.default_rng(0).uniform(0, 1, 5), device=dev)
small = lambda: us.integrate(c, lambda s: torch.exp(c.to_cartesian_dot(s, kk)), False, 16, xp=torch, device=dev, dtype=torch.float64)
small(); torch.cuda.synchronize()
side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    small()
torch.cuda.current_stream().wait_stream(side)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    res = small()
def many(fn, reps=200):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps * 1e6
out["integrate_small|standard:4|n16|fused"] = {"eager_us": many(small), "graph_replay_us": many(g.replay), "value": res.item()}
print(json.dumps(out, indent=1))
