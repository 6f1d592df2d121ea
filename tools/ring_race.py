"""Hunt for nondeterminism in the ring kernels: run from_cartesian / to_cartesian repeatedly on fixed
inputs and compare every result with the first, bit for bit; print which outputs and which points differ.
    python tools/ring_race.py random:32:0 f32 64000000 30 [knob=value,...]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ultrasphere_b200 as us
from ultrasphere_b200 import _native
from tests.helpers import make

spec, dt, n, reps = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
if len(sys.argv) > 5:
    for kv in sys.argv[5].split(","):
        k, v = kv.split("=")
        assert _native.lib().us_set_tuning(k.encode(), int(v)) == 0
tdt = torch.float32 if dt == "f32" else torch.float64
c = make(us, spec)
x = torch.rand((c.c_ndim, n), device="cuda", dtype=tdt, generator=torch.Generator(device="cuda").manual_seed(1234)).mul_(2).sub_(1)
s0 = c.from_cartesian(x)
y0 = c.to_cartesian(s0, as_array=True)
torch.cuda.synchronize()
bad_from = bad_to = 0
for rep in range(reps):
    s = c.from_cartesian(x)
    for i, k in enumerate(s0):
        d = (s[k] != s0[k]) & ~(s[k].isnan() & s0[k].isnan())
        m = int(d.sum())
        if m:
            idx = d.nonzero().flatten()
            bad_from += m
            print(f"rep {rep} from[{k}] (output {i}): {m} differ, first {idx[:3].tolist()} last {int(idx[-1])}; mod 2048: {[int(j) % 2048 for j in idx[:3].tolist()]}", flush=True)
    y = c.to_cartesian(s0, as_array=True)
    d = (y != y0).any(0)
    m = int(d.sum())
    if m:
        idx = d.nonzero().flatten()
        rows = (y[:, idx[0]] != y0[:, idx[0]]).nonzero().flatten().tolist()
        bad_to += m
        print(f"rep {rep} to: {m} points differ, first {idx[:3].tolist()} last {int(idx[-1])}; mod 2048: {[int(j) % 2048 for j in idx[:3].tolist()]}; rows wrong at the first: {rows[:8]}... ({len(rows)} of {c.c_ndim})", flush=True)
print("ring_race", spec, dt, n, "reps", reps, sys.argv[5] if len(sys.argv) > 5 else "", "bad from", bad_from, "bad to", bad_to, flush=True)
