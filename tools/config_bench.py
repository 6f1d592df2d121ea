"""Throughput of the secondary BASELINE configs (C1, C3, C4) at their full sizes: per-kernel
time, points/s and fraction of the HBM roofline.  Prints one JSON object."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ultrasphere_b200 as us
from tests.helpers import make

dev = torch.device("cuda:0")
CONFIGS = [("spherical", torch.float64, 100_000), ("spherical", torch.float64, 300_000), ("spherical", torch.float64, 1_000_000), ("spherical", torch.float64, 4_000_000),
           ("standard:9", torch.float32, 1_000_000), ("standard:9", torch.float32, 4_000_000), ("spherical", torch.float64, 64_000_000), ("hopf:3", torch.float64, 128_000_000),
           ("standard_prime:8", torch.float64, 128_000_000), ("random:32:0", torch.float32, 64_000_000), ("standard:9", torch.float64, 128_000_000)]
only = os.environ.get("US_ONLY")
peak = 6536.4
if os.path.isfile("MEASURED_PEAKS.json"):
    peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"]

def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(reps):
        e0.record(); fn(); e1.record(); e1.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best

def t_graph(fn, calls=20, reps=5):
    """Per-call device time with the host launch cost removed: `calls` launches captured in one CUDA graph."""
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(calls):
            keep = fn()
    return t(g.replay, reps) / calls

out = {}
for spec, dt, n in CONFIGS:
    if only and only not in spec:
        continue
    c = make(us, spec)
    x = torch.rand((c.c_ndim, n), device=dev, dtype=dt) * 2 - 1
    s = c.from_cartesian(x)
    bpp = 2 * c.c_ndim * x.element_size()
    tf = t(lambda: c.from_cartesian(x)); tt = t(lambda: c.to_cartesian(s, as_array=True))
    out[f"{spec}|{str(dt)[6:]}|{n}"] = {
        "from_ms": tf, "to_ms": tt, "from_gpts": n / tf / 1e6, "to_gpts": n / tt / 1e6,
        "from_frac_hbm": n * bpp / tf / 1e6 / peak, "to_frac_hbm": n * bpp / tt / 1e6 / peak, "bytes_per_point": bpp,
    }
    if n <= 4_000_000:  # launch-bound sizes: the same calls replayed from a CUDA graph
        gf = t_graph(lambda: c.from_cartesian(x)); gt = t_graph(lambda: c.to_cartesian(s, as_array=True))
        out[f"{spec}|{str(dt)[6:]}|{n}"].update({"graph_from_ms": gf, "graph_to_ms": gt,
            "graph_from_frac_hbm": n * bpp / gf / 1e6 / peak, "graph_to_frac_hbm": n * bpp / gt / 1e6 / peak})
    del x, s
    torch.cuda.empty_cache()
# FP64 / FP32 FMA peaks of this GPU (dependent-free DFMA / FFMA loops on every SM): the denominator of
# the pipe-bound float64 kernels, which MEASURED_PEAKS.json does not carry
import ctypes
from ultrasphere_b200 import _native
for code, name in ((1, "fp64"), (0, "fp32")):
    tf = ctypes.c_double()
    _native.check(_native.lib().us_probe_fma_peak(code, 4096, ctypes.byref(tf)), "probe")
    out[f"measured_fma_peak_{name}_tflops"] = tf.value
print(json.dumps(out, indent=1))
