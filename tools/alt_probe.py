"""Does a kernel's time depend on what ran right before it?  Same buffers throughout (C ABI row
entry points, preallocated outputs), per-call CUDA events:
  same      from, from, from, ...           (and to, to, to, ...)
  alternate from, to, from, to, ...         (what a round-trip loop does)
  alloc     alternate, through the public API (fresh outputs from the caching allocator each call)
python tools/alt_probe.py standard:9 f32 256000000"""
import ctypes as C, os, statistics, sys
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ultrasphere_b200 as us  # noqa: E402
from ultrasphere_b200 import _native  # noqa: E402
from ultrasphere_b200._plan import plan_for  # noqa: E402
from tests.helpers import make  # noqa: E402

spec, dt, n = sys.argv[1], sys.argv[2], int(sys.argv[3])
reps = int(os.environ.get("REPS", "40"))
tdt = torch.float32 if dt == "f32" else torch.float64
code, esz = (0, 4) if dt == "f32" else (1, 8)
c = make(us, spec)
h = _native.lib()
plan = plan_for(c).plan
nn = c.s_ndim
x = torch.rand((c.c_ndim, n), device="cuda", dtype=tdt).mul_(2).sub_(1)
out_from = torch.empty((nn + 1, n), device="cuda", dtype=tdt)
out_to = torch.empty((c.c_ndim, n), device="cuda", dtype=tdt)
row = n * esz
stream = torch.cuda.current_stream().cuda_stream
ang = (C.c_void_p * nn)(*[out_from.data_ptr() + (1 + i) * row for i in range(nn)])


def f():
    assert h.us_from_cartesian_rows(C.byref(plan), code, n, x.data_ptr(), row, out_from.data_ptr(), row, stream) == 0


def t():
    assert h.us_to_cartesian_rows(C.byref(plan), code, n, ang, out_from.data_ptr(), out_to.data_ptr(), row, stream) == 0


def timed(seq):
    evs = []
    for name, fn in seq:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        evs.append((name, e0, e1))
    torch.cuda.synchronize()
    out = {}
    for name, e0, e1 in evs:
        out.setdefault(name, []).append(e0.elapsed_time(e1))
    return {k: (round(statistics.median(v), 4), round(min(v), 4)) for k, v in out.items()}


f(); t(); torch.cuda.synchronize()
print(spec, dt, n, "(median, best) ms")
print("same      ", timed([("from", f)] * reps), timed([("to", t)] * reps))
print("alternate ", timed([("from", f), ("to", t)] * reps))
print("same      ", timed([("from", f)] * reps), timed([("to", t)] * reps))
print("alternate ", timed([("from", f), ("to", t)] * reps))
state = {}
def fa():
    state["s"] = c.from_cartesian(x)
def ta():
    state["b"] = c.to_cartesian(state["s"], as_array=True)
fa(); ta(); torch.cuda.synchronize()
print("alloc     ", timed([("from", fa), ("to", ta)] * reps))
print("alloc same", timed([("from", fa)] * reps), timed([("to", ta)] * reps))
