"""Stress one transform config at a large batch: every from_cartesian r against the float64 norm torch
computes, every round trip against the input; prints where mismatches sit (tile pattern).
    python tools/stress.py standard:9 f32 256000000 [reps]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ultrasphere_b200 as us
from tests.helpers import make

spec, dt, n = sys.argv[1], sys.argv[2], int(sys.argv[3])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
tdt = torch.float32 if dt == "f32" else torch.float64
eps = torch.finfo(tdt).eps
c = make(us, spec)
x = torch.rand((c.c_ndim, n), device="cuda", dtype=tdt, generator=torch.Generator(device="cuda").manual_seed(1234)).mul_(2).sub_(1)
bad_total = 0
chunk = 1 << 24
for rep in range(reps):
    s = c.from_cartesian(x)
    y = c.to_cartesian(s, as_array=True)
    torch.cuda.synchronize()
    for lo in range(0, n, chunk):
        sl = slice(lo, min(n, lo + chunk))
        r64 = x[:, sl].double().square().sum(0).sqrt()
        bad_r = ((s["r"][sl].double() - r64).abs() > c.c_ndim * eps * r64)
        bad_y = ((y[:, sl] - x[:, sl]).abs().amax(0) > 16 * eps * r64.to(tdt))
        for name, bad in (("r", bad_r), ("roundtrip", bad_y)):
            k = int(bad.sum())
            if k:
                idx = bad.nonzero().flatten()[:6] + lo
                bad_total += k
                print(f"rep {rep} {name}: {k} bad points in [{lo}, {sl.stop}); first at {idx.tolist()} (tile of 512: {[int(i) // 512 for i in idx.tolist()]})", flush=True)
print("stress", spec, dt, n, "reps", reps, "bad", bad_total, flush=True)
