"""Paired A/B INSIDE one process: variants alternate call by call on the same buffers, so the slow
drift of the HBM-bound kernels (several percent over tens of milliseconds on one box) hits all of
them alike.  A variant is a library build plus tuning knobs:

    python tools/ab_inproc.py standard:9 f32 256000000 ab_libs/r01x.so ab_libs/v7.so ab_libs/v7.so:tile_groups=4 ab_libs/v7.so:tile_groups=5,tile_bufs=9

Calls go straight through the C ABI (us_from_cartesian_rows / us_to_cartesian_rows) into
preallocated outputs; per-call CUDA-event times; prints median / mean / best per variant and
direction, with the fraction of the measured HBM peak."""
import ctypes as C, json, os, statistics, sys
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ultrasphere_b200 as us  # noqa: E402  (host-side tree construction and plan listing only)
from ultrasphere_b200 import _native  # noqa: E402
from ultrasphere_b200._plan import preorder_listing  # noqa: E402
from tests.helpers import make  # noqa: E402

spec, dt, n = sys.argv[1], sys.argv[2], int(sys.argv[3])
variants = sys.argv[4:]
reps = int(os.environ.get("AB_REPS", "15"))
tdt = torch.float32 if dt == "f32" else torch.float64
code = 0 if dt == "f32" else 1
esz = 4 if dt == "f32" else 8
c = make(us, spec)
nodes, kinds, leaves = preorder_listing(c)
s_index = {node: i for i, node in enumerate(c.s_nodes)}
c_index = {node: i for i, node in enumerate(c.c_nodes)}
nn = len(nodes)

libs = {}
plans = {}
for v in variants:
    path = v.split(":")[0]
    if path in libs:
        continue
    h = C.CDLL(os.path.abspath(path))
    for name, (res, args) in _native._SIGNATURES.items():
        try:
            fn = getattr(h, name)
        except AttributeError:
            continue
        fn.restype, fn.argtypes = res, args
    plan = _native.UsPlan()
    rc = h.us_plan_compile(nn, (C.c_uint8 * nn)(*kinds), (C.c_int32 * nn)(*[s_index[x] for x in nodes]),
                           (C.c_int32 * (nn + 1))(*[c_index[x] for x in leaves]), C.byref(plan))
    assert rc == 0, (path, rc)
    libs[path], plans[path] = h, plan

x = torch.rand((c.c_ndim, n), device="cuda", dtype=tdt).mul_(2).sub_(1)
out_from = torch.empty((nn + 1, n), device="cuda", dtype=tdt)
out_to = torch.empty((c.c_ndim, n), device="cuda", dtype=tdt)
row = n * esz
stream = torch.cuda.current_stream().cuda_stream
ang = (C.c_void_p * nn)(*[out_from.data_ptr() + (1 + i) * row for i in range(nn)])


def apply(v):
    path, *knobs = v.split(":")
    h = libs[path]
    if hasattr(h, "us_set_tuning"):
        for name in ("tile_groups", "tile_bufs", "tile_u"):
            h.us_set_tuning(name.encode(), 0)
        h.us_set_tuning(b"tile_max_rows", 13)
        h.us_set_tuning(b"ring_ctas", 0)
        h.us_set_tuning(b"ring_u", 0)  # libraries older than the knob answer US_EINVAL: ignored
        h.us_set_tuning(b"ring_slots", 0)
        if knobs:
            for kv in knobs[0].split(","):
                k, val = kv.split("=")
                assert h.us_set_tuning(k.encode(), int(val)) == 0
    return h, plans[path]


def call_from(v):
    h, plan = apply(v)
    rc = h.us_from_cartesian_rows(C.byref(plan), code, n, x.data_ptr(), row, out_from.data_ptr(), row, stream)
    assert rc == 0, (v, rc, h.us_last_error_string())


def call_to(v):
    h, plan = apply(v)
    rc = h.us_to_cartesian_rows(C.byref(plan), code, n, ang, out_from.data_ptr(), out_to.data_ptr(), row, stream)
    assert rc == 0, (v, rc, h.us_last_error_string())


times = {(v, d): [] for v in variants for d in ("from", "to")}
block = int(os.environ.get("AB_BLOCK", "0"))  # > 0: sustained mode — blocks of this many back-to-back calls per variant
# untimed calls of the variant right before each timed block: under the 1000 W cap the SM clock settles
# to the variant's own power draw only after several hundred milliseconds (tools/clocks_probe.py), so
# energy differences between variants show only when each one runs alone for that long
settle = int(os.environ.get("AB_SETTLE", "0"))
if block > 0:
    for direction, fn in (("from", call_from), ("to", call_to)):
        for v in variants:
            fn(v)
        torch.cuda.synchronize()
        for rep in range(reps):
            order = variants if rep % 2 == 0 else variants[::-1]
            for v in order:
                for _ in range(settle):
                    fn(v)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(block):
                    fn(v)
                e1.record()
                e1.synchronize()
                times[(v, direction)].append(e0.elapsed_time(e1) / block)
        if direction == "from":
            call_from(variants[-1])
# AB_BENCH=1: what bench.py (and the driver) does — after an idle second, 5 untimed and 20 timed round
# trips (from, to, from, to, ...) back to back, per-kernel events; variants alternate per repetition
if os.environ.get("AB_BENCH"):
    import time
    for rep in range(reps):
        order = variants if rep % 2 == 0 else variants[::-1]
        for v in order:
            torch.cuda.synchronize()
            time.sleep(float(os.environ.get("AB_IDLE", "1.0")))
            for _ in range(5):
                call_from(v); call_to(v)
            evs = []
            for _ in range(20):
                e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
                e[0].record(); call_from(v); e[1].record(); call_to(v); e[2].record()
                evs.append(e)
            torch.cuda.synchronize()
            times[(v, "from")].append(sum(e[0].elapsed_time(e[1]) for e in evs) / 20)
            times[(v, "to")].append(sum(e[1].elapsed_time(e[2]) for e in evs) / 20)
    block = -1
for direction, fn in ((("from", call_from), ("to", call_to)) if block == 0 else ()):
    for v in variants:
        fn(v)
    torch.cuda.synchronize()
    for rep in range(reps):
        order = variants if rep % 2 == 0 else variants[::-1]
        for v in order:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(v); e1.record(); e1.synchronize()
            times[(v, direction)].append(e0.elapsed_time(e1))
    if direction == "from":
        call_from(variants[-1])  # the angles to_cartesian reads
err = float((out_to - x).abs().max())
peak = 6536.4
pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.isfile(pk):
    peak = json.load(open(pk))["hbm_gbs"]
gb = n * 2 * c.c_ndim * esz / 1e9
print(f"{spec} {dt} n={n} reps={reps} {'blocks of ' + str(block) + ' back-to-back calls' if block > 0 else ('bench pattern: idle, 5 + 20 round trips' if block < 0 else 'isolated calls')} round-trip max err {err:.2e}")
print("| variant | from median ms (frac) | from best | to median ms (frac) | to best |\n|---|---|---|---|---|")
for v in variants:
    f, t = times[(v, "from")], times[(v, "to")]
    fm, tm = statistics.median(f), statistics.median(t)
    print(f"| {v} | {fm:.4f} ({gb / fm * 1e3 / peak:.3f}) | {min(f):.4f} | {tm:.4f} ({gb / tm * 1e3 / peak:.3f}) | {min(t):.4f} |")
