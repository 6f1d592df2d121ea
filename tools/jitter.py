"""Where does the step-to-step spread of the headline kernels come from?  Times from_cartesian on
create_standard(9) float32, 256 M points, (a) through the public API (a fresh 10 GB output every
call), (b) straight through the C ABI into ONE preallocated output, (c) the same into a second
allocation — 24 calls each, per-call CUDA-event times."""
import ctypes as C, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ultrasphere_b200 as us
from ultrasphere_b200 import _native
from ultrasphere_b200._plan import plan_for

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256_000_000
c = us.create_standard(9)
ct = plan_for(c)
x = torch.rand((10, n), device="cuda").mul_(2).sub_(1)
lib = _native.lib()
stream = torch.cuda.current_stream().cuda_stream

def series(fn, k=24):
    for _ in range(3): fn()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(k + 1)]
    torch.cuda.synchronize()
    ev[0].record()
    for i in range(k):
        fn(); ev[i + 1].record()
    torch.cuda.synchronize()
    return [round(ev[i].elapsed_time(ev[i + 1]), 3) for i in range(k)]

out = {}
out["api_fresh_output"] = series(lambda: c.from_cartesian(x))
row = n * 4
for name in ("abi_fixed_output_A", "abi_fixed_output_B", "abi_fixed_output_C"):
    buf = torch.empty((10, n), device="cuda")
    out[name] = series(lambda: lib.us_from_cartesian_rows(ct.plan_ref, 0, n, x.data_ptr(), row, buf.data_ptr(), row, stream))
    out[name + "_ptr_mod_2MB"] = buf.data_ptr() % (1 << 21)
    keep = buf  # hold A while B is allocated so that B is a different block
    if name.endswith("A"): hold = buf
s = c.from_cartesian(x)
sb = torch.stack([s[k] for k in c.s_nodes])
ang = (C.c_void_p * 9)(*[sb[i].data_ptr() for i in range(9)])
for name in ("to_fixed_output_A", "to_fixed_output_B"):
    buf = torch.empty((10, n), device="cuda")
    out[name] = series(lambda: lib.us_to_cartesian_rows(ct.plan_ref, 0, n, ang, s["r"].data_ptr(), buf.data_ptr(), row, stream))
    hold2 = buf
print(json.dumps(out, indent=1))
