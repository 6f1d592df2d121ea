"""Per-call host overhead of the public API (tiny batches, GPU time negligible)."""
import os, sys, time, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ultrasphere_b200 as us
from tests.helpers import make

out = {}
for spec in ("spherical", "standard:9", "random:32:0"):
    c = make(us, spec)
    for n in (1024, 1 << 20):
        x = torch.rand((c.c_ndim, n), device="cuda", dtype=torch.float32) * 2 - 1
        s = c.from_cartesian(x)
        for name, fn in (("from", lambda: c.from_cartesian(x)), ("to", lambda: c.to_cartesian(s)), ("to_array", lambda: c.to_cartesian(s, as_array=True))):
            for _ in range(20):
                fn()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            reps = 300
            for _ in range(reps):
                fn()
            t_issue = (time.perf_counter() - t0) / reps * 1e6
            torch.cuda.synchronize()
            t_total = (time.perf_counter() - t0) / reps * 1e6
            out[f"{spec}|n={n}|{name}"] = {"host_us_per_call": round(t_issue, 1), "wall_us_per_call": round(t_total, 1)}
print(json.dumps(out, indent=1))
