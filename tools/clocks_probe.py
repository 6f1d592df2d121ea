"""Do the clocks move under the transform kernels?  Runs one direction of one config back to back for
a while (default 1.5 s) while a thread samples NVML every 2 ms: SM / memory clock, power, temperatures
and the throttle-reason mask; prints per-call time early vs late in the run next to the trace.
    python tools/clocks_probe.py hopf:3 f64 128000000 [seconds]"""
import json, os, sys, threading, time
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ultrasphere_b200 as us  # noqa: E402
from tests.helpers import make  # noqa: E402
import pynvml  # noqa: E402

spec, dt, n = sys.argv[1], sys.argv[2], int(sys.argv[3])
seconds = float(sys.argv[4]) if len(sys.argv) > 4 else 1.5
tdt = torch.float32 if dt == "f32" else torch.float64
c = make(us, spec)
x = torch.rand((c.c_ndim, n), device="cuda", dtype=tdt) * 2 - 1
s = c.from_cartesian(x)
y = c.to_cartesian(s, as_array=True)
torch.cuda.synchronize()
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(int(os.environ.get("CUDA_VISIBLE_DEVICES", "0").split(",")[0]))
samples, stop = [], threading.Event()


def pump():
    while not stop.is_set():
        row = {"t": time.perf_counter(), "sm": pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM),
               "mem": pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_MEM), "w": pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0,
               "temp": pynvml.nvmlDeviceGetTemperature(h, pynvml.NVML_TEMPERATURE_GPU)}
        try:
            row["mask"] = int(pynvml.nvmlDeviceGetCurrentClocksEventReasons(h))
        except Exception:
            row["mask"] = int(pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h))
        samples.append(row)
        time.sleep(0.002)


out = {"spec": spec, "dtype": dt, "n": n}
for direction, fn in (("from", lambda: c.from_cartesian(x)), ("to", lambda: c.to_cartesian(s, as_array=True))):
    time.sleep(0.5)  # let the GPU idle between the two runs
    samples.clear()
    stop.clear()
    th = threading.Thread(target=pump, daemon=True)
    th.start()
    time.sleep(0.05)
    ev, t0 = [], time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            ev.append((e0, e1))
        torch.cuda.synchronize()
    stop.set()
    th.join()
    ms = [a.elapsed_time(b) for a, b in ev]
    k = max(1, len(ms) // 10)
    step = max(1, len(samples) // 12)
    out[direction] = {
        "calls": len(ms), "ms_first_tenth": sorted(ms[:k])[k // 2], "ms_last_tenth": sorted(ms[-k:])[k // 2], "ms_min": min(ms), "ms_median": sorted(ms)[len(ms) // 2],
        "sm_mhz": [r["sm"] for r in samples[::step]], "mem_mhz": sorted({r["mem"] for r in samples}), "power_w": [round(r["w"]) for r in samples[::step]],
        "temp_c": [r["temp"] for r in samples[::step]], "reason_masks": sorted({hex(r["mask"]) for r in samples}),
    }
print(json.dumps(out))
