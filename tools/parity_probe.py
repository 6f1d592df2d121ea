"""Development probe (GPU box): error statistics of the CUDA path against the NumPy oracle and
quick per-kernel timings.  Writes gpurun_out/parity_probe.json.  Not part of the product."""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ultrasphere_b200 as us  # noqa: E402
from oracle import vks_oracle as vo  # noqa: E402
from tests.helpers import make  # noqa: E402

SPECS = ["spherical", "standard:9", "hopf:3", "standard_prime:8", "random:32:0", "standard:4"]


def stats(got: np.ndarray, want: np.ndarray, scale: np.ndarray | None = None) -> dict:
    d = vo.ulp_distance(got, want)
    out = {"max_ulp": float(d.max()), "p999_ulp": float(np.quantile(d, 0.999)), "frac_gt4": float((d > 4).mean())}
    g, w = got.astype(np.float64), want.astype(np.float64)
    with np.errstate(all="ignore"):
        rel = np.abs(g - w) / np.abs(w)
    out["max_rel"] = float(np.nanmax(np.where(np.isfinite(rel), rel, 0)))
    if scale is not None:
        eps = np.finfo(got.dtype).eps
        out["max_abs_over_r_eps"] = float(np.max(np.abs(g - w) / scale.astype(np.float64)) / eps)
    return out


def timeit(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(reps):
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def main() -> None:
    dev = torch.device("cuda:0")
    n = int(os.environ.get("PROBE_N", 1 << 20))
    res: dict = {"n": n, "gpu": torch.cuda.get_device_name(0)}
    for spec in SPECS:
        c, t = make(us, spec), vo.make(spec)
        for dname, dt, tdt in (("f32", np.float32, torch.float32), ("f64", np.float64, torch.float64)):
            rng = np.random.default_rng(0)
            x = rng.uniform(-1, 1, (c.c_ndim, n)).astype(dt)
            xd = torch.from_numpy(x).to(dev)
            want = vo.from_cartesian(t, x)
            got = c.from_cartesian(xd)
            entry: dict = {}
            entry["key_order_ok"] = list(got) == list(want)
            entry["r"] = stats(got["r"].cpu().numpy(), want["r"])
            ang_g = np.stack([got[k].cpu().numpy() for k in c.s_nodes])
            ang_w = np.stack([want[k] for k in c.s_nodes])
            entry["angles"] = stats(ang_g, ang_w)
            entry["angles_abs_eps_pi"] = float(np.max(np.abs(ang_g.astype(np.float64) - ang_w)) / (np.finfo(dt).eps * np.pi))
            # to_cartesian on the ORACLE's angles so that only this kernel's error shows
            sph_d = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in want.items()}
            back = c.to_cartesian(sph_d, as_array=True).cpu().numpy()
            wantx = vo.to_cartesian(t, want, as_array=True)
            entry["leaves"] = stats(back, wantx, want["r"][None])
            # full GPU round trip vs the input
            rt = c.to_cartesian(got, as_array=True).cpu().numpy()
            entry["roundtrip_max_abs"] = float(np.max(np.abs(rt.astype(np.float64) - x)))
            entry["roundtrip_max_abs_over_r_eps"] = float(
                np.max(np.abs(rt.astype(np.float64) - x) / want["r"][None]) / np.finfo(dt).eps
            )
            # timings with the data resident (1M points: launch + L2 effects dominate; indicative only)
            big = 1 << 24
            xb = torch.rand((c.c_ndim, big), device=dev, dtype=tdt) * 2 - 1
            sb = c.from_cartesian(xb)
            t_from = timeit(lambda: c.from_cartesian(xb))
            t_to = timeit(lambda: c.to_cartesian(sb, as_array=True))
            bpp = 2 * c.c_ndim * np.dtype(dt).itemsize
            entry["from_gpts"] = big / t_from / 1e6
            entry["to_gpts"] = big / t_to / 1e6
            entry["from_gbs"] = big * bpp / t_from / 1e6
            entry["to_gbs"] = big * bpp / t_to / 1e6
            del xb, sb
            res[f"{spec}|{dname}"] = entry
            print(spec, dname, json.dumps(entry), flush=True)
    # integrals
    for spec, nq in (("standard:4", 24), ("hopf:2", 16), ("spherical", 32), ("expr:cbaa", 12)):
        c, t = make(us, spec), vo.make(spec)
        k = np.random.default_rng(0).uniform(0, 1, c.c_ndim)
        kd = torch.from_numpy(k).to(dev)
        ref = vo.integrate(t, lambda s: np.exp(np.einsum("v,v...->...", k, vo.to_cartesian(t, s, as_array=True))), False, nq)
        for slab in (None, 0, 1 << 16):
            got = us.integrate(
                c, lambda s: torch.exp(torch.einsum("v,v...->...", kd, c.to_cartesian(s, as_array=True))),
                False, nq, xp=torch, device=dev, dtype=torch.float64, slab_bytes=slab,
            ).item()
            res[f"int|{spec}|{slab}"] = {"got": got, "ref": float(ref), "rel": abs(got - float(ref)) / abs(float(ref))}
            print("integral", spec, slab, res[f"int|{spec}|{slab}"], flush=True)
    from ultrasphere_b200 import _native
    import ctypes

    for code, name in ((1, "fp64"), (0, "fp32")):
        tf = ctypes.c_double()
        _native.check(_native.lib().us_probe_fma_peak(code, 4096, ctypes.byref(tf)), "probe")
        res[f"fma_peak_{name}_tflops"] = tf.value
        print("fma peak", name, tf.value)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "parity_probe.json"), "w") as fh:
        json.dump(res, fh, indent=1)


if __name__ == "__main__":
    main()
