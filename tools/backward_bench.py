"""Throughput of the backward kernels (autograd through the transforms): create_standard(9) float32
and create_hopf(3) float64, 16 M points.  Prints one JSON object."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ultrasphere_b200 as us
from tests.helpers import make

dev = torch.device("cuda:0")
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(reps):
        e0.record(); fn(); e1.record(); e1.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best
out = {}
for spec, dt in (("standard:9", torch.float32), ("hopf:3", torch.float64)):
    c = make(us, spec)
    n = 16_000_000
    x = (torch.rand((c.c_ndim, n), device=dev, dtype=dt) * 2 - 1).requires_grad_()
    s = c.from_cartesian(x)
    g = torch.rand((1 + c.s_ndim, n), device=dev, dtype=dt)
    stacked = torch.stack([s["r"]] + [s[k] for k in c.s_nodes])
    fwd_from = t(lambda: c.from_cartesian(x))
    bwd_from = t(lambda: torch.autograd.grad(stacked, x, g, retain_graph=True))
    sd = {k: v.detach().requires_grad_() for k, v in s.items()}
    y = c.to_cartesian(sd, as_array=True)
    gy = torch.rand_like(y)
    fwd_to = t(lambda: c.to_cartesian(sd, as_array=True))
    bwd_to = t(lambda: torch.autograd.grad(y, list(sd.values()), gy, retain_graph=True))
    esz = x.element_size()
    out[f"{spec}|{str(dt)[6:]}|{n}"] = {
        "from_forward_autograd_ms": fwd_from, "from_backward_ms": bwd_from, "to_forward_autograd_ms": fwd_to, "to_backward_ms": bwd_to,
        "from_backward_gpts": n / bwd_from / 1e6, "to_backward_gpts": n / bwd_to / 1e6,
        "note": "backward_ms includes autograd's own stack/unbind copies of the cotangents",
    }
    del x, s, sd, y, gy, g, stacked
    torch.cuda.empty_cache()
print(json.dumps(out, indent=1))
