"""Summarise extra ncu reports (gpurun_out/prof_rNN_*.ncu-rep) into profiles/rNN_ncu_extra.md."""
import csv, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
R = sys.argv[1] if len(sys.argv) > 1 else "r01"
REPORTS = [
    (f"prof_{R}_hopf3_f64", "create_hopf(3) float64, 32 M points (128 B/point/direction: 4.10 GB algorithmic per launch)", 32e6 * 128),
    (f"prof_{R}_ring", "create_random(32, default_rng(0)) float32, 16 M points, ring kernels (264 B/point/direction: 4.22 GB)", 16e6 * 264),
    (f"prof_{R}_c5", "integrate C5 pieces: broadcast grid kernel (6.79 GB written), row kernel of the fused functional (1.36 GB written) and TMA-streamed reduction (1.36 GB read)", None),
]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"]
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
out = [f"# {R} — `ncu --set full --clock-control none` on the float64, wide-tree and quadrature kernels", "",
       "Each report: the plain command exited 0 first, then `ncu --set full --clock-control none --import-source on -k regex:<kernel> -s <warm-up> -c 2`.", ""]
for name, what, alg in REPORTS:
    path = os.path.join(ROOT, "gpurun_out", name + ".ncu-rep")
    if not os.path.isfile(path):
        continue
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    out += [f"## {what}", ""]
    for r in rows[2:]:
        out += [f"### `{r[hdr.index('Kernel Name')]}`", "", "| metric | value | unit |", "|---|---|---|"]
        for k in KEYS:
            if k in hdr:
                out.append(f"| {k} | {r[hdr.index(k)]} | {units[hdr.index(k)]} |")
        rd = float(r[hdr.index("dram__bytes_read.sum")]) * SCALE[units[hdr.index("dram__bytes_read.sum")]]
        wr = float(r[hdr.index("dram__bytes_write.sum")]) * SCALE[units[hdr.index("dram__bytes_write.sum")]]
        dur = float(r[hdr.index("gpu__time_duration.sum")]) * {"ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}[units[hdr.index("gpu__time_duration.sum")]]
        line = f"| **DRAM traffic** | {(rd + wr) / 1e9:.3f} GB in {dur * 1e3:.3f} ms = {(rd + wr) / dur / 1e9:.0f} GB/s under the profiler |"
        if alg:
            line += f" algorithmic {alg / 1e9:.3f} GB, ratio {(rd + wr) / alg:.3f} |"
        else:
            line += " |"
        out += [line, ""]
open(os.path.join(ROOT, "profiles", f"{R}_ncu_extra.md"), "w").write("\n".join(out) + "\n")
print("\n".join(l for l in out if l.startswith("###") or "DRAM traffic" in l or "issue_active" in l or "fp64" in l))
