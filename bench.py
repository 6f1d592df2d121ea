#!/usr/bin/env python
"""Benchmark of the hot path.  Headline: create_standard(9) (10-D `bbbbbbbba`) float32 round trip
(from_cartesian → to_cartesian), BASELINE.json configs[1]; every other BASELINE config is a leg of
the same line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

One process per GPU (torchrun for N>1).  A step is one round trip over one batch of
POINTS_PER_GPU synthetic points resident in HBM.  Rank 0 prints ONE JSON line.

  value      whole-job round-trip points/s, inputs resident in HBM, CUDA-event timed, max over ranks
  roofline   the dominant kernel's algorithmic HBM bytes / its event-timed duration / measured peak
  e2e        the same round trip through the public API from pinned HOST buffers, copies included
  cpu_baseline  the reference's own code (oracle/_ref, else the NumPy port) on this box's host, one
             core, bounded sample
  configs    the other BASELINE configs at their full sizes (per GPU): C1 create_spherical() float64
             1 M, C3 create_hopf(3) and create_standard_prime(8) float64 128 M, C4 create_random(32)
             float32 64 M — per direction ms, points/s, fraction of the HBM roofline and, for
             float64, of the FP64 pipe (DFMA peak measured in the run by us_probe_fma_peak)
  quadrature BASELINE configs[4]: us.integrate on create_standard(4), n = 96 (170 M nodes, float64),
             grid sharded over the ranks with one all-reduce; nodes/s for the reference-form
             integrand and for the fused functional (outside the timed steps)
  multi_rank_check  (N > 1) sharded integrals against the same integrals computed unsharded

``--impl reference`` times the UNMODIFIED reference (the four modules `oracle/make_ref.py` copies
into the git-ignored `oracle/_ref/`, loaded by `oracle/refload.py`) on all host cores; the NumPy
port in `oracle/` stands in only where that copy is absent.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

S_NDIM = 9
C_NDIM = 10
POINTS_PER_GPU = 256_000_000
E2E_POINTS = 32_000_000
E2E_CHUNK = 4_000_000
CPU_SAMPLE = 1 << 23
METRIC = "round-trip points/s (create_standard(9) float32: from_cartesian then to_cartesian)"
UNIT = "points/s"
BYTES_PER_POINT_ONE_WAY = 2 * C_NDIM * 4  # (n_in + n_out) * sizeof(float32)
L2_BYTES = 126 << 20

# (key, tree spec, dtype name, points per GPU, BASELINE.json config it is)
LEGS = (
    ("C1", "spherical", "f64", 1_000_000, "configs[0]: create_spherical() `ba` 3-D float64 round trip, 1 M points"),
    ("C3_hopf3", "hopf:3", "f64", 128_000_000, "configs[2]: create_hopf(3) `ccaacaa` 8-D float64 round trip, 128 M points"),
    ("C3_standard_prime8", "standard_prime:8", "f64", 128_000_000, "configs[2]: create_standard_prime(8) 9-D float64 round trip, 128 M points"),
    ("C4", "random:32:0", "f32", 64_000_000, "configs[3]: create_random(32, rng=default_rng(0)) 33-D float32 round trip, 64 M points"),
)


def parse() -> argparse.Namespace:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--points", type=int, default=POINTS_PER_GPU, help="points per GPU (default: the BASELINE size)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-quadrature", action="store_true", help="skip the integrate() leg (BASELINE configs[4])")
    ap.add_argument("--no-configs", action="store_true", help="skip the C1 / C3 / C4 legs")
    ap.add_argument("--leg-scale", type=float, default=1.0, help="scale the C3 / C4 leg sizes (tests use a small fraction)")
    ap.add_argument("--no-sustained", action="store_true", help="skip the power-settled (about 1 s of back-to-back calls) timings")
    ap.add_argument("--sustain-seconds", type=float, default=0.9, help="untimed back-to-back calls before each sustained timing")
    return ap.parse_args()


def make_tree(us, spec: str):
    """Coordinate object for a spec string, with any module exposing the reference's create_*."""
    import numpy as np

    head, *rest = spec.split(":")
    if head == "random":
        return us.create_random(int(rest[0]), rng=np.random.default_rng(int(rest[1])))
    if head in ("polar", "spherical"):
        return getattr(us, f"create_{head}")()
    return getattr(us, f"create_{head}")(int(rest[0]))


# ----------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock, power and throttle reasons sampled DURING the timed region (NVML from a thread,
    ~2 ms period; falls back to one nvidia-smi loop if NVML cannot be loaded)."""

    REASONS = (
        ("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4),
        ("hw_power_brake_slowdown", 0x80),
    )

    def __init__(self, index: int) -> None:
        self.index = index
        self.samples: list[tuple[float, float, int, float]] = []  # (MHz, W, reason mask, host time)
        self.mem_mhz: list[tuple[float, float]] = []  # (memory MHz, host time)
        self.window: tuple[float, float] | None = None
        self.max_mhz: float | None = None
        self._stop = threading.Event()
        self._thread: threading.Thread | None = None
        self._mode = "none"

    def _physical_index(self) -> int:
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            try:
                return int(vis.split(",")[self.index])
            except (ValueError, IndexError):
                pass
        return self.index

    def start(self) -> None:
        try:
            import pynvml

            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index())
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))

            def pump() -> None:
                while not self._stop.is_set():
                    try:
                        mhz = float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                        watts = pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0
                        try:
                            mask = int(pynvml.nvmlDeviceGetCurrentClocksEventReasons(h))
                        except Exception:
                            mask = int(pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                        now = time.perf_counter()
                        self.samples.append((mhz, watts, mask, now))
                        self.mem_mhz.append((float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_MEM)), now))
                    except Exception:
                        pass
                    time.sleep(0.001)

            self._mode = "nvml"
        except Exception:
            fields = "clocks.sm,clocks.max.sm,power.draw"
            try:
                proc = subprocess.Popen(
                    ["nvidia-smi", f"--id={self.index}", f"--query-gpu={fields}", "--format=csv,noheader,nounits", "-lms", "20"],
                    stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True,
                )
            except OSError:
                return
            self._proc = proc

            def pump() -> None:
                assert proc.stdout is not None
                for line in proc.stdout:
                    cells = [c.strip() for c in line.split(",")]
                    try:
                        self.samples.append((float(cells[0]), float(cells[2]), 0, time.perf_counter()))
                        self.max_mhz = float(cells[1])
                    except (ValueError, IndexError):
                        continue

            self._mode = "nvidia-smi"
        self._thread = threading.Thread(target=pump, daemon=True)
        self._thread.start()

    def stop(self) -> dict:
        self._stop.set()
        if self._mode == "nvidia-smi":
            self._proc.terminate()
        if self._thread is not None:
            self._thread.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["no samples"], "source": self._mode}
        # samples inside the timed region; the sampler also runs through the warm-up steps (same
        # kernels, same load), which are used only if the timed region was too short to be sampled
        picked = self.samples
        scope = "warm-up + timed region"
        if self.window is not None:
            inside = [x for x in self.samples if self.window[0] <= x[3] <= self.window[1]]
            if len(inside) >= 3:
                picked, scope = inside, "timed region"
        peak_w = max(w for _, w, _, _ in picked)
        hot = sorted(m for m, w, _, _ in picked if w >= 0.5 * peak_w) or sorted(m for m, _, _, _ in picked)
        mask = 0
        for _, _, m, _ in picked:
            mask |= m
        out = {
            "sm_mhz": hot[len(hot) // 2], "sm_max_mhz": self.max_mhz, "power_w_max": peak_w, "samples": len(picked),
            "scope": scope, "reasons": [name for name, bit in self.REASONS if mask & bit], "source": self._mode,
        }
        # how the clocks moved THROUGH the timed region (ten evenly spaced samples): a kernel that is
        # timed for tens of milliseconds settles below its first steps when the power cap bites
        t_lo, t_hi = picked[0][3], picked[-1][3]
        mem = [m for m, t in self.mem_mhz if t_lo <= t <= t_hi]
        if mem:
            out["mem_mhz"] = sorted(mem)[len(mem) // 2]
            out["mem_mhz_min"] = min(mem)
        step = max(1, len(picked) // 10)
        out["trace"] = {"sm_mhz": [x[0] for x in picked[::step]][:10], "power_w": [round(x[1]) for x in picked[::step]][:10],
                        "sw_power_cap": [int(bool(x[2] & 0x4)) for x in picked[::step]][:10]}
        return out


# ----------------------------------------------------------------------------------------------
# CPU legs: the unmodified reference (oracle/_ref or the mount) where it can be loaded, else the port
# ----------------------------------------------------------------------------------------------
class CpuImpl:
    """The reference's CPU path behind one face: `kind` "reference" = its own unmodified modules
    (oracle/refload.py), "port" = the NumPy restatement in oracle/vks_oracle.py."""

    def __init__(self, prefer_reference: bool = True) -> None:
        self.kind = "port"
        self.note = "NumPy port of the reference's path (oracle/vks_oracle.py); oracle/_ref is absent"
        if prefer_reference:
            try:
                from oracle import refload

                if refload.available():
                    self._ref = refload.load()
                    self.kind = "reference"
                    self.note = f"unmodified ultrasphere modules from {refload.source_dir()} (oracle/refload.py), NumPy arrays"
            except Exception as exc:  # a missing dependency of the reference on this box
                self.note = f"NumPy port (loading the reference failed: {type(exc).__name__}: {exc})"
        if self.kind == "port":
            from oracle import vks_oracle

            self._vo = vks_oracle

    def tree(self, spec: str):
        if self.kind == "reference":
            return make_tree(self._ref.us, spec)
        return self._vo.make(spec)

    def round_trip(self, tree, x) -> None:
        if self.kind == "reference":
            tree.to_cartesian(tree.from_cartesian(x), as_array=True)
        else:
            self._vo.to_cartesian(tree, self._vo.from_cartesian(tree, x), as_array=True)

    def integrate_exp(self, n: int) -> tuple[float, float]:
        """(value, seconds) of integrate(create_standard(4), exp(k.x), n) in float64."""
        import numpy as np

        k5 = np.random.default_rng(0).uniform(0, 1, 5)
        t4 = self.tree("standard:4")
        if self.kind == "reference":
            f = lambda s_: np.exp(np.einsum("v,v...->...", k5, t4.to_cartesian(s_, as_array=True)))  # noqa: E731
            run = lambda m: self._ref.us.integrate(t4, f, False, m, xp=np)  # noqa: E731
        else:
            f = lambda s_: np.exp(np.einsum("v,v...->...", k5, self._vo.to_cartesian(t4, s_, as_array=True)))  # noqa: E731
            run = lambda m: self._vo.integrate(t4, f, False, m)  # noqa: E731
        run(8)
        t0 = time.perf_counter()
        got = float(run(n))
        return got, time.perf_counter() - t0


def cpu_baseline_single_core(sample: int) -> dict:
    import numpy as np

    impl = CpuImpl()
    tree = impl.tree(f"standard:{S_NDIM}")
    x = np.random.default_rng(0).uniform(-1, 1, (C_NDIM, sample)).astype(np.float32)
    impl.round_trip(tree, x[:, : sample // 8])
    best = float("inf")
    for _ in range(2):
        t0 = time.perf_counter()
        impl.round_trip(tree, x)
        best = min(best, time.perf_counter() - t0)
    return {
        "value": sample / best,
        "unit": UNIT,
        "cores": 1,
        "kind": impl.kind,
        "sample": f"{sample} points of the same workload, best of 2, NumPy ufuncs are single-threaded (host has {os.cpu_count()} cpus)",
        "what": impl.note,
    }


def run_reference(args: argparse.Namespace, emit) -> None:
    """The reference's CPU implementation on every host core: the batch is cut into one slice per
    thread and each thread calls the reference's own from_cartesian / to_cartesian on its slice
    (NumPy releases the GIL inside ufuncs; the reference itself is single-threaded)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from concurrent.futures import ThreadPoolExecutor

    import numpy as np

    impl = CpuImpl()
    cores = os.cpu_count() or 1
    tree = impl.tree(f"standard:{S_NDIM}")
    sample = min(args.points, CPU_SAMPLE)
    x = np.random.default_rng(0).uniform(-1, 1, (C_NDIM, sample)).astype(np.float32)
    bounds = np.linspace(0, sample, cores + 1).astype(int)
    parts = [np.ascontiguousarray(x[:, lo:hi]) for lo, hi in zip(bounds[:-1], bounds[1:]) if hi > lo]
    with ThreadPoolExecutor(max_workers=cores) as pool:
        def step() -> None:
            list(pool.map(lambda part: impl.round_trip(tree, part), parts))

        for _ in range(args.warmup):
            step()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step()
        dt = time.perf_counter() - t0
    value = sample * args.steps / dt
    # the stock call: one thread, the whole sample in one from_cartesian / to_cartesian pair
    t0 = time.perf_counter()
    impl.round_trip(tree, x)
    single = sample / (time.perf_counter() - t0)
    desc = f"{sample} points per step split over {len(parts)} threads"
    # quadrature leg on the host: the same integrand on a bounded grid (n = 40: 5.12 M nodes, ~0.5 s)
    quadrature = None
    if not args.no_quadrature:
        nq = 40
        got, qdt = impl.integrate_exp(nq)
        nodes = nq * nq * nq * 2 * nq
        quadrature = {"workload": "integrate(create_standard(4), exp(k.x), n=40) float64 on the host (bounded sample of the n=96 grid)",
                      "reference_form": {"ms_per_integral": qdt * 1e3, "nodes_per_s": nodes / qdt, "value": got}, "cores": 1, "kind": impl.kind}
    port = None
    if impl.kind == "reference" and not args.no_cpu_baseline:
        # second key: the NumPy port of the same path, threaded the same way (the r01 reference arm)
        pimpl = CpuImpl(prefer_reference=False)
        ptree = pimpl.tree(f"standard:{S_NDIM}")
        with ThreadPoolExecutor(max_workers=cores) as pool:
            list(pool.map(lambda part: pimpl.round_trip(ptree, part), parts))
            t0 = time.perf_counter()
            list(pool.map(lambda part: pimpl.round_trip(ptree, part), parts))
            port = {"value": sample / (time.perf_counter() - t0), "unit": UNIT, "cores": cores, "kind": "port", "sample": desc + ", one step"}
    emit(json.dumps({
        "impl": "reference", "quadrature": quadrature,
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "create_standard(9) bbbbbbbba float32 round trip (BASELINE.json configs[1])", "points_per_step": sample,
                   "runs_on": f"host CPU, {impl.note}"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": impl.kind, "sample": desc,
                         "single_thread": {"value": single, "unit": UNIT, "cores": 1, "sample": f"{sample} points, one stock call pair"}},
        "port": port,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ----------------------------------------------------------------------------------------------
# B200 arm
# ----------------------------------------------------------------------------------------------
def fp64_instr_per_point(s_ndim: int, c_ndim: int) -> dict:
    """FP64-pipe instructions (DFMA/DMUL/DADD/DSETP) per point of the float64 kernels, counted in
    the source (us_common.cuh) and confirmed by ncu's sm__inst_executed_pipe_fp64:
    from_cartesian 34 per node (atan2 23 + sqrt 8 + norm 3) plus r (c_ndim squares, c_ndim - 1 adds,
    sqrt ~10); to_cartesian 32 per node (sincos 30 + two products)."""
    return {"from": 34 * s_ndim + 2 * c_ndim + 9, "to": 32 * s_ndim}


def run_b200(args: argparse.Namespace, emit) -> None:
    import ctypes

    import torch
    import torch.distributed as dist

    import ultrasphere_b200 as us
    from ultrasphere_b200 import _native
    from ultrasphere_b200._host import HostPipeline

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 arm has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n = args.points
    c = us.create_standard(S_NDIM)
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    x = torch.rand((C_NDIM, n), device=dev, dtype=torch.float32, generator=gen).mul_(2).sub_(1)

    def barrier() -> None:
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(values: list[float]) -> list[float]:
        t = torch.tensor(values, device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(v) for v in t.tolist()]

    def sustained_pair(tree, x_in, sph_in, n_pts: int, bytes_per_point: int) -> dict:
        """Both directions after the clocks have settled.  These kernels draw the board's full 1000 W:
        within a few hundred milliseconds of back-to-back launches the power cap pulls the SM clock
        from 1965 MHz to 1.5-1.65 GHz (profiles/r02_clocks_*.json), so a timing of 10-20 steps right
        after an idle phase (the numbers above) is a burst figure.  Here every direction runs untimed
        for --sustain-seconds first, then ~0.3 s are timed with NVML sampled next to it."""
        out = {"how": f"{args.sustain_seconds} s of back-to-back calls untimed, then about 0.3 s timed (CUDA events, max over ranks); clocks = NVML on rank 0 during the timed part"}
        for direction, fn in (("from_cartesian", lambda: tree.from_cartesian(x_in)), ("to_cartesian", lambda: tree.to_cartesian(sph_in, as_array=True))):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            per = max((time.perf_counter() - t0) / 3, 1e-5)
            n_settle, n_timed = max(5, int(args.sustain_seconds / per)), max(5, int(0.3 / per))
            barrier()
            for _ in range(n_settle):
                fn()
            probe = ClockSampler(local)
            if rank == 0:
                probe.start()
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            w0 = time.perf_counter()
            s0.record()
            for _ in range(n_timed):
                fn()
            s1.record()
            torch.cuda.synchronize()
            probe.window = (w0, time.perf_counter())
            ck = probe.stop() if rank == 0 else None
            ms = max_over_ranks([s0.elapsed_time(s1) / n_timed])[0]
            gbs = n_pts * bytes_per_point / (ms * 1e-3) / 1e9
            out[direction] = {"ms": ms, "calls_timed": n_timed, "achieved_gbs": gbs, "frac": gbs / peak,
                              "clocks": None if ck is None else {k: ck.get(k) for k in ("sm_mhz", "sm_max_mhz", "power_w_max", "reasons", "mem_mhz")}}
        out["value"] = world * n_pts / ((out["from_cartesian"]["ms"] + out["to_cartesian"]["ms"]) * 1e-3)
        out["unit"] = UNIT
        return out

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(peaks_path):
        with open(peaks_path) as fh:
            peak, peak_src = float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    warmup = max(args.warmup, 3)
    sph = None
    for _ in range(warmup):
        sph = c.from_cartesian(x)
        back = c.to_cartesian(sph, as_array=True)
    # parity spot check inside the bench run: the round trip must reproduce the input
    err = ((back[:, : 1 << 20] - x[:, : 1 << 20]).abs().amax(0) / sph["r"][: 1 << 20]).max().item()
    if not err <= 8 * 1.1920929e-07:
        raise SystemExit(f"bench.py: round-trip error {err} exceeds 8 eps — refusing to report a number")
    del back
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    barrier()
    host_t0 = time.perf_counter()
    t_begin = torch.cuda.Event(enable_timing=True)
    t_end = torch.cuda.Event(enable_timing=True)
    t_begin.record()
    for k in range(args.steps):
        ev[k][0].record()
        sph = c.from_cartesian(x)
        ev[k][1].record()
        back = c.to_cartesian(sph, as_array=True)
        ev[k][2].record()
    t_end.record()
    barrier()
    sampler.window = (host_t0, time.perf_counter())
    clocks = sampler.stop() if rank == 0 else None
    total_ms = t_begin.elapsed_time(t_end)
    from_each = [e[0].elapsed_time(e[1]) for e in ev]
    to_each = [e[1].elapsed_time(e[2]) for e in ev]
    from_ms = sum(from_each) / args.steps
    to_ms = sum(to_each) / args.steps
    total_ms, from_ms, to_ms = max_over_ranks([total_ms, from_ms, to_ms])
    value = world * n * args.steps / (total_ms * 1e-3)
    del back
    headline_sustained = None if args.no_sustained else sustained_pair(c, x, sph, n, 2 * C_NDIM * 4)
    del sph, x
    torch.cuda.empty_cache()

    # The roofline's denominator (MEASURED_PEAKS.json) is a BURST figure: the best of ten isolated
    # 2 GiB copies.  The timed region above runs ~150 ms of back-to-back HBM-bound launches, which the
    # part does not sustain at burst rate (sw_power_cap shows in the clock record).  So the same
    # copy is timed here the way the steps are — back to back for about as long — and reported next
    # to the burst figure: `frac` stays achieved / burst peak, `frac_of_sustained_copy` says how far
    # the kernel is from what a plain copy sustains on this box in this run.
    sustained = None
    try:
        src = torch.empty(1 << 30, dtype=torch.bfloat16, device=dev).normal_()
        dst = torch.empty_like(src)
        for _ in range(5):
            dst.copy_(src)
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        per_copy = 2 * src.numel() * 2 / (peak * 1e9)
        settle_copies = 0 if args.no_sustained else int(args.sustain_seconds / per_copy)
        copies = 200 if args.no_sustained else max(200, int(0.3 / per_copy))
        for _ in range(settle_copies):
            dst.copy_(src)
        probe = ClockSampler(local)
        if rank == 0:
            probe.start()
        torch.cuda.synchronize()
        w0 = time.perf_counter()
        c0.record()
        for _ in range(copies):
            dst.copy_(src)
        c1.record()
        torch.cuda.synchronize()
        probe.window = (w0, time.perf_counter())
        copy_clocks = probe.stop() if rank == 0 else None
        one = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        best = float("inf")
        for _ in range(10):
            torch.cuda.synchronize()
            time.sleep(0.02)
            one[0].record()
            dst.copy_(src)
            one[1].record()
            torch.cuda.synchronize()
            best = min(best, one[0].elapsed_time(one[1]))
        gb = 2 * src.numel() * 2 / 1e9
        sus = -max_over_ranks([-(gb * copies / (c0.elapsed_time(c1) * 1e-3))])[0]  # the slowest rank's
        sustained = {"sustained_copy_gbs": sus, "burst_copy_gbs": gb / (best * 1e-3), "copies_back_to_back": copies,
                     "seconds": c0.elapsed_time(c1) * 1e-3, "settle_copies_untimed": settle_copies,
                     "clocks": None if copy_clocks is None else {k: copy_clocks.get(k) for k in ("sm_mhz", "sm_max_mhz", "power_w_max", "reasons", "mem_mhz")},
                     "how": "torch b.copy_(a) over 1 Gi bf16 elements (read + write bytes): copies back to back after --sustain-seconds of the same untimed vs the best of 10 isolated ones, same run, after the timed steps"}
        del src, dst
        torch.cuda.empty_cache()
    except Exception as exc:
        sustained = {"error": f"{type(exc).__name__}: {exc}"}

    # ---- the other BASELINE configs (C1, C3, C4), same bar: full size per GPU, events, max over ranks ----
    configs = None
    if not args.no_configs:
        configs = {}
        fma = {}
        for code, name in ((_native.US_F64, "fp64"), (_native.US_F32, "fp32")):
            tf = ctypes.c_double()
            _native.check(_native.lib().us_probe_fma_peak(code, 4096, ctypes.byref(tf)), "us_probe_fma_peak")
            fma[name] = tf.value
        fma["fp64"], fma["fp32"] = max_over_ranks([-fma["fp64"], -fma["fp32"]])  # min over ranks
        fma = {k: -v for k, v in fma.items()}
        configs["measured_fma_peak_tflops"] = fma
        # the same probe once the power cap has settled the clocks: ~1 s of back-to-back DFMA launches,
        # the last call's figure (the denominator for the float64 legs' sustained FP64-pipe fractions)
        fma_sustained = None
        if not args.no_sustained:
            tf = ctypes.c_double()
            t_end_probe = time.perf_counter() + args.sustain_seconds + 0.3
            while time.perf_counter() < t_end_probe:
                _native.check(_native.lib().us_probe_fma_peak(_native.US_F64, 32768, ctypes.byref(tf)), "us_probe_fma_peak")
            fma_sustained = -max_over_ranks([-tf.value])[0]
            configs["measured_fma_peak_tflops_sustained"] = {"fp64": fma_sustained, "how": f"us_probe_fma_peak(float64) called back to back for {args.sustain_seconds + 0.3:.1f} s, last call (best of its 3 launches)"}
        leg_steps = max(3, min(args.steps, 10))
        for key, spec, dname, n_leg, what in LEGS:
            tdt = torch.float32 if dname == "f32" else torch.float64
            if key != "C1":
                n_leg = max(1 << 20, int(n_leg * args.leg_scale))
            ct = make_tree(us, spec)
            esz = 4 if dname == "f32" else 8
            bpp = 2 * ct.c_ndim * esz
            # inputs larger than L2 (126 MB): C1's 24 MB batch is rotated through enough distinct copies
            set_bytes = ct.c_ndim * n_leg * esz
            n_sets = 1 if set_bytes > 2 * L2_BYTES else -(-3 * L2_BYTES // set_bytes)
            g = torch.Generator(device=dev).manual_seed(4321 + rank)
            xs = [torch.rand((ct.c_ndim, n_leg), device=dev, dtype=tdt, generator=g).mul_(2).sub_(1) for _ in range(n_sets)]
            sphs = [ct.from_cartesian(xi) for xi in xs]
            back = None
            for i in range(max(3, n_sets)):
                sphs[i % n_sets] = ct.from_cartesian(xs[i % n_sets])
                back = ct.to_cartesian(sphs[i % n_sets], as_array=True)
            eps = 1.1920929e-07 if dname == "f32" else 2.220446049250313e-16
            j = (max(3, n_sets) - 1) % n_sets
            m = min(n_leg, 1 << 20)
            rt_err = ((back[:, :m] - xs[j][:, :m]).abs().amax(0) / sphs[j]["r"][:m]).max().item()
            if not rt_err <= 16 * eps:
                configs[key] = {"workload": what, "error": f"round-trip error {rt_err} exceeds 16 eps: not reported"}
                continue
            calls = leg_steps * n_sets
            e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            barrier()
            e[0].record()
            for i in range(calls):
                sphs[i % n_sets] = ct.from_cartesian(xs[i % n_sets])
            e[1].record()
            for i in range(calls):
                back = ct.to_cartesian(sphs[i % n_sets], as_array=True)
            e[2].record()
            barrier()
            f_ms, t_ms = max_over_ranks([e[0].elapsed_time(e[1]) / calls, e[1].elapsed_time(e[2]) / calls])
            leg = {
                "workload": what, "dtype": dname, "points_per_gpu": n_leg, "calls_timed": calls,
                "l2": (f"{n_sets} distinct input sets rotated ({n_sets * set_bytes >> 20} MB > L2)" if n_sets > 1 else "inputs exceed the 126 MB L2"),
                "from_cartesian_ms": f_ms, "to_cartesian_ms": t_ms,
                "value": world * n_leg / ((f_ms + t_ms) * 1e-3), "unit": UNIT,
                "from_points_per_s": world * n_leg / (f_ms * 1e-3), "to_points_per_s": world * n_leg / (t_ms * 1e-3),
                "round_trip_err_over_r": rt_err,
                "roofline": {
                    "bound": "hbm", "unit": "GB/s", "peak": peak, "algorithmic_bytes_per_point": bpp,
                    "from_cartesian": {"achieved": n_leg * bpp / (f_ms * 1e-3) / 1e9, "frac": n_leg * bpp / (f_ms * 1e-3) / 1e9 / peak},
                    "to_cartesian": {"achieved": n_leg * bpp / (t_ms * 1e-3) / 1e9, "frac": n_leg * bpp / (t_ms * 1e-3) / 1e9 / peak},
                },
            }
            if dname == "f64":
                ipp = fp64_instr_per_point(ct.s_ndim, ct.c_ndim)
                pipe_peak = fma["fp64"] / 2 * 1e12  # FP64-pipe instructions/s: one DFMA = 2 flop
                leg["fp64_pipe"] = {
                    "unit": "lane-instructions/s", "peak": pipe_peak, "peak_source": "us_probe_fma_peak (dependent-free DFMA on every SM, this run)",
                    "from_cartesian": {"instr_per_point": ipp["from"], "frac": ipp["from"] * n_leg / (f_ms * 1e-3) / pipe_peak},
                    "to_cartesian": {"instr_per_point": ipp["to"], "frac": ipp["to"] * n_leg / (t_ms * 1e-3) / pipe_peak},
                }
            if key == "C1":
                # the same calls without the host's launch cost: 20 calls per CUDA graph, replayed
                def graph_ms(fn) -> float:
                    side = torch.cuda.Stream()
                    side.wait_stream(torch.cuda.current_stream())
                    with torch.cuda.stream(side):
                        fn(0)
                    torch.cuda.current_stream().wait_stream(side)
                    gr = torch.cuda.CUDAGraph()
                    keep = []
                    with torch.cuda.graph(gr):
                        for i in range(20):
                            keep.append(fn(i))
                    gr.replay()
                    torch.cuda.synchronize()
                    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    g0.record()
                    for _ in range(5):
                        gr.replay()
                    g1.record()
                    torch.cuda.synchronize()
                    return g0.elapsed_time(g1) / 100

                gf = graph_ms(lambda i: ct.from_cartesian(xs[i % n_sets]))
                gt = graph_ms(lambda i: ct.to_cartesian(sphs[i % n_sets], as_array=True))
                gf, gt = max_over_ranks([gf, gt])
                leg["eager_note"] = "eager calls: at 1 M points the host's per-call work, not the kernel, sets the time"
                leg["graph_replay"] = {"from_cartesian_ms": gf, "to_cartesian_ms": gt,
                                       "from_frac_hbm": n_leg * bpp / (gf * 1e-3) / 1e9 / peak, "to_frac_hbm": n_leg * bpp / (gt * 1e-3) / 1e9 / peak}
            if not args.no_sustained and key != "C1":
                leg["sustained"] = sustained_pair(ct, xs[0], sphs[0], n_leg, bpp)
                if dname == "f64" and fma_sustained:
                    ipp = fp64_instr_per_point(ct.s_ndim, ct.c_ndim)
                    pk = fma_sustained / 2 * 1e12
                    leg["sustained"]["fp64_pipe"] = {
                        "peak": pk, "peak_source": "measured_fma_peak_tflops_sustained",
                        "from_cartesian_frac": ipp["from"] * n_leg / (leg["sustained"]["from_cartesian"]["ms"] * 1e-3) / pk,
                        "to_cartesian_frac": ipp["to"] * n_leg / (leg["sustained"]["to_cartesian"]["ms"] * 1e-3) / pk,
                    }
            configs[key] = leg
            del xs, sphs, back
            torch.cuda.empty_cache()

    # ---- end to end from host buffers (per rank, same tree, E2E_POINTS points) -----------------
    e2e = None
    if not args.no_e2e:
        ne = min(E2E_POINTS, n)
        hx = {k: torch.empty(ne, dtype=torch.float32).uniform_(-1, 1).pin_memory() for k in c.c_nodes}
        hs = {k: torch.empty(ne, dtype=torch.float32).pin_memory() for k in ["r"] + list(c.s_nodes)}
        hb = {k: torch.empty(ne, dtype=torch.float32).pin_memory() for k in c.c_nodes}
        pipe = HostPipeline(c, device=dev, chunk_points=E2E_CHUNK)
        for _ in range(2):
            pipe.round_trip(hx, hs, hb)
        esteps = max(3, min(args.steps, 10))
        barrier()
        t0 = time.perf_counter()
        for _ in range(esteps):
            pipe.round_trip(hx, hs, hb)
        barrier()
        dt = max_over_ranks([time.perf_counter() - t0])[0]
        # variant for information: only the round trip's final result travels back (the spherical
        # coordinates stay in HBM like the intermediate they are in the device-resident number)
        for _ in range(2):
            pipe.round_trip(hx, None, hb)
        barrier()
        t0 = time.perf_counter()
        for _ in range(esteps):
            pipe.round_trip(hx, None, hb)
        barrier()
        dt_final = max_over_ranks([time.perf_counter() - t0])[0]
        d2h_final = pipe.bytes_d2h
        pipe.round_trip(hx, hs, hb)  # restore the full-result byte counters for the report below
        # ceiling of the host links for THIS traffic pattern, measured in the same run on every rank
        # at once: raw pinned copies, H2D and D2H concurrently on two streams, in the ratio the
        # round trip moves them (40 B/point in, 80 B/point out), nothing else on the GPU
        link = None
        try:
            mb = 1 << 20
            h_in = torch.empty(160 * mb, dtype=torch.uint8).pin_memory()
            h_out = torch.empty(320 * mb, dtype=torch.uint8).pin_memory()
            d_in = torch.empty(160 * mb, dtype=torch.uint8, device=dev)
            d_out = torch.empty(320 * mb, dtype=torch.uint8, device=dev)
            s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

            def raw(reps: int) -> None:
                for _ in range(reps):
                    with torch.cuda.stream(s_in):
                        d_in.copy_(h_in, non_blocking=True)
                    with torch.cuda.stream(s_out):
                        h_out.copy_(d_out, non_blocking=True)
                s_in.synchronize()
                s_out.synchronize()

            raw(2)
            barrier()
            t0 = time.perf_counter()
            raw(6)
            barrier()
            ldt = max_over_ranks([time.perf_counter() - t0])[0]
            h2d_gbs, d2h_gbs = world * 6 * 160 * mb / ldt / 1e9, world * 6 * 320 * mb / ldt / 1e9
            link = {"h2d_gbs": h2d_gbs, "d2h_gbs": d2h_gbs, "ranks_copying_at_once": world,
                    "round_trip_ceiling_points_per_s": d2h_gbs * 1e9 / (2 * C_NDIM * 4),
                    "how": "raw pinned cudaMemcpyAsync, 160 MB host->device and 320 MB device->host concurrently on two streams per rank, 6 repetitions, wall clock, max over ranks"}
            del h_in, h_out, d_in, d_out
        except Exception as exc:  # a box that cannot pin another 480 MB: the ceiling is simply not reported
            link = {"error": f"{type(exc).__name__}: {exc}"}
        bad = max((hb[k][:100000] - hx[k][:100000]).abs().max().item() for k in c.c_nodes)
        if not bad <= 4e-6:
            raise SystemExit(f"bench.py: e2e round-trip error {bad}")
        e2e = {
            "value": world * ne * esteps / dt,
            "unit": UNIT,
            "h2d_bytes_per_step": pipe.bytes_h2d,
            "d2h_bytes_per_step": pipe.bytes_d2h,
            "points_per_step_per_gpu": ne,
            "steps": esteps,
            "launches_per_step": pipe.launches,
            "host_link_gbs": {"h2d": world * pipe.bytes_h2d * esteps / dt / 1e9, "d2h": world * pipe.bytes_d2h * esteps / dt / 1e9,
                              "note": "aggregate over all ranks during the timed steps: the box's host links bound this number, not the kernels"},
            "host_link_ceiling": link,
            "frac_of_link_ceiling": (world * ne * esteps / dt) / link["round_trip_ceiling_points_per_s"] if link and "round_trip_ceiling_points_per_s" in link else None,
            "final_result_only": {"value": world * ne * esteps / dt_final, "unit": UNIT, "d2h_bytes_per_step": d2h_final,
                                  "note": "same pipeline, only the reconstructed Cartesian points are copied back"},
            "how": "pinned host dict-of-arrays -> HostPipeline.round_trip (chunked H2D / from_cartesian / to_cartesian / D2H of spherical and Cartesian results on 3 streams), wall clock around the calls, max over ranks",
        }
        del hx, hs, hb, pipe

    # ---- quadrature leg: the second half of the metric (BASELINE.json configs[4]) ----------------
    # us.integrate on create_standard(4), n = 96 (169,869,312 nodes), float64, f = exp(k.x); the
    # leading grid axis is sharded over the ranks (strong scaling) with one NCCL all-reduce of the
    # scalar.  Two integrands: the reference's own form (materialises the Cartesian grid, then
    # torch.einsum) and the fused functional.  CUDA events, max over ranks.
    quadrature = None
    multi_rank_check = None
    if not args.no_quadrature:
        torch.cuda.empty_cache()
        import numpy as np

        from scipy.special import iv

        c4 = us.create_standard(4)
        k5 = np.random.default_rng(0).uniform(0, 1, 5)
        kd = torch.from_numpy(k5).to(dev)
        kn = float(np.linalg.norm(k5))
        want = float((2 * np.pi) ** 2.5 * kn ** (1 - 2.5) * iv(1.5, kn))  # integral of exp(k.x) over S^4, closed form
        nq = 96
        nodes = nq * nq * nq * 2 * nq
        group = dist.group.WORLD if world > 1 else None
        forms = {
            "reference_form": lambda sph_: torch.exp(torch.einsum("v,v...->...", kd, c4.to_cartesian(sph_, as_array=True))),
            "fused_functional": lambda sph_: torch.exp(c4.to_cartesian_dot(sph_, kd)),
        }
        quadrature = {"workload": "us.integrate(create_standard(4), exp(k.x), n=96) float64, 169,869,312 nodes",
                      "scaling": "strong (leading grid axis sharded over the ranks, one all-reduce of the scalar)",
                      "analytic": want}
        for name, f in forms.items():
            run = lambda: us.integrate(c4, f, False, nq, xp=torch, device=dev, dtype=torch.float64, group=group)  # noqa: E731
            for _ in range(3):
                got = run()
            barrier()
            q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 10
            q0.record()
            for _ in range(reps):
                got = run()
            q1.record()
            barrier()
            ms = max_over_ranks([q0.elapsed_time(q1) / reps])[0]
            rel = abs(got.item() - want) / want
            if not rel <= 1e-12:  # the same all-reduced value on every rank: no number for this leg, the line still prints
                quadrature[name] = {"error": f"integral off by {rel} relative: not reported", "rel_err_vs_analytic": rel}
                continue
            quadrature[name] = {"ms_per_integral": ms, "nodes_per_s": nodes / (ms * 1e-3), "rel_err_vs_analytic": rel}
        # the fused form once more, captured in ONE CUDA graph (f's kernels, the contraction and the
        # NCCL all-reduce of the sharded partial sums): what is left is device time
        try:
            captured = us.IntegralGraph(c4, forms["fused_functional"], False, nq, xp=torch, device=dev, dtype=torch.float64, group=group)
            for _ in range(3):
                got = captured()
            barrier()
            q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 20
            q0.record()
            for _ in range(reps):
                got = captured()
            q1.record()
            barrier()
            ms = max_over_ranks([q0.elapsed_time(q1) / reps])[0]
            rel = abs(got.item() - want) / want
            if rel <= 1e-12:
                quadrature["fused_functional_graph"] = {"ms_per_integral": ms, "nodes_per_s": nodes / (ms * 1e-3), "rel_err_vs_analytic": rel,
                                                        "how": "us.IntegralGraph: the whole sharded integral, all-reduce included, replayed as one CUDA graph"}
            else:
                quadrature["fused_functional_graph"] = {"error": f"integral off by {rel} relative: not reported", "rel_err_vs_analytic": rel}
            del captured
        except Exception as exc:  # a box whose NCCL cannot be captured: the eager legs above stand
            quadrature["fused_functional_graph"] = {"error": f"{type(exc).__name__}: {exc}"}
        if world > 1:
            # the check of tests/test_gpu_multi.py under THIS run's ranks: every sharded integral
            # (two trees, both integrand forms, slabbed) against the same integral unsharded
            worst = 0.0
            cases = 0
            for ct, nn_ in ((us.create_standard(4), 24), (us.create_hopf(2), 17)):
                kk = torch.from_numpy(np.random.default_rng(0).uniform(0, 1, ct.c_ndim)).to(dev)
                plain = lambda s_, ct=ct, kk=kk: torch.exp(torch.einsum("v,v...->...", kk, ct.to_cartesian(s_, as_array=True)))  # noqa: E731
                fused = lambda s_, ct=ct, kk=kk: torch.exp(ct.to_cartesian_dot(s_, kk))  # noqa: E731
                kw = dict(xp=torch, device=dev, dtype=torch.float64)
                single = us.integrate(ct, plain, False, nn_, **kw).item()
                for fn, extra in ((plain, {}), (fused, {"slab_bytes": 1 << 16})):
                    got = us.integrate(ct, fn, False, nn_, group=group, **extra, **kw).item()
                    worst = max(worst, abs(got - single) / abs(single))
                    cases += 1
            worst = max_over_ranks([worst])[0]
            multi_rank_check = {"ranks": world, "cases": cases, "max_rel_diff_sharded_vs_unsharded": worst, "ok": bool(worst <= 1e-12),
                                "what": "integrate(group=WORLD) on create_standard(4) n=24 and create_hopf(2) n=17, both integrand forms, against group=None on the same GPU"}
        torch.cuda.empty_cache()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    dom, dom_ms = ("us::from_cartesian", from_ms) if from_ms >= to_ms else ("us::to_cartesian", to_ms)
    alg_bytes = BYTES_PER_POINT_ONE_WAY * n
    achieved = alg_bytes / (dom_ms * 1e-3) / 1e9
    # DRAM traffic of the dominant kernel: dram__bytes_read+write from the committed `ncu --set full`
    # capture (profiles/traffic_rNN.json), scaled to this launch's size if the capture's differs
    traffic, traffic_src = None, None
    tfiles = sorted(f for f in os.listdir(os.path.join(ROOT, "profiles")) if f.startswith("traffic_r") and f.endswith(".json")) if os.path.isdir(os.path.join(ROOT, "profiles")) else []
    if tfiles:
        with open(os.path.join(ROOT, "profiles", tfiles[-1])) as fh:
            rec = json.load(fh).get(dom)
        if rec:
            traffic = rec["ratio"] * alg_bytes
            traffic_src = f"profiles/{tfiles[-1]}: measured/algorithmic = {rec['ratio']:.4f} at {rec.get('points', 64_000_000) / 1e6:.0f} M points, scaled to this launch"
    roofline = {
        "bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
        "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes,
        "ms_per_launch": dom_ms,
        "other_kernel": {"from_cartesian_ms": from_ms, "to_cartesian_ms": to_ms,
                         "from_cartesian_gbs": alg_bytes / (from_ms * 1e-3) / 1e9, "to_cartesian_gbs": alg_bytes / (to_ms * 1e-3) / 1e9,
                         "from_cartesian_frac": alg_bytes / (from_ms * 1e-3) / 1e9 / peak, "to_cartesian_frac": alg_bytes / (to_ms * 1e-3) / 1e9 / peak},
        "per_step_ms_rank0": {"from_cartesian": [round(v, 4) for v in from_each], "to_cartesian": [round(v, 4) for v in to_each]},
        "copy_in_this_run": sustained,
        "sustained": headline_sustained,
        "frac_of_sustained_copy": (achieved / sustained["sustained_copy_gbs"]) if sustained and "sustained_copy_gbs" in sustained else None,
    }
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warmup,
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {
            "workload": "create_standard(9) bbbbbbbba float32 round trip (BASELINE.json configs[1])",
            "points_per_gpu": n, "global_points": world * n, "parallelism": f"batch sharded over {world} GPU(s), no collective",
            "l2": "inputs (10.24 GB per array set) exceed the 126 MB L2; no flush needed",
            "inputs": "U(-1,1) per leaf, torch.Generator(cuda).manual_seed(1234+rank), resident in HBM",
        },
        "roofline": roofline, "e2e": e2e, "gpu_launches": 2 * args.steps, "clocks": clocks,
        "configs": configs, "quadrature": quadrature, "multi_rank_check": multi_rank_check,
    }
    if not args.no_cpu_baseline:
        if world == 1:
            line["cpu_baseline"] = cpu_baseline_single_core(min(CPU_SAMPLE, n))
        else:
            line["cpu_baseline"] = {"value": None, "note": "timed on rank 0 at N=1 only (the contract): see the --gpus 1 line of the same scaling run"}
    emit(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main() -> None:
    args = parse()
    # The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version
    # banner on the first communicator), so file descriptor 1 is pointed at stderr for the whole
    # run and the JSON line goes to the saved, real stdout at the very end.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    lines: list[str] = []
    emit = lines.append
    try:
        if args.impl == "reference":
            run_reference(args, emit)
        else:
            run_b200(args, emit)
    finally:
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        os.close(real_stdout)
    for line in lines:
        sys.stdout.write(line + "\n")
    sys.stdout.flush()


if __name__ == "__main__":
    main()
