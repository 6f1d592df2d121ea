"""Turn what tools/evidence_run.sh left under gpurun_out/ into the committed summaries under profiles/
(run in the build container; needs the ncu CLI to read the reports and cuobjdump for the SASS).

    python tools/make_profiles_r02.py [r02]
"""
import collections, csv, json, os, shutil, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
R = sys.argv[1] if len(sys.argv) > 1 else "r02"
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)

COPIES = [
    (f"bench_{R}.json", f"{R}_bench.json"), (f"bench_{R}_reference.json", f"{R}_bench_reference.json"),
    (f"launches_{R}.csv", f"{R}_launches.csv"), ("parity_report.json", f"{R}_parity_report.json"),
    (f"{R}_host_overhead.json", f"{R}_host_overhead.json"), (f"{R}_jitter.json", f"{R}_jitter.json"),
    (f"{R}_checked_smoke.log", f"{R}_checked_smoke.log"), (f"{R}_gpu_tests.log", f"{R}_gpu_tests.log"),
    (f"{R}_clocks_std9_f32.json", f"{R}_clocks_std9_f32.json"), (f"{R}_clocks_hopf3_f64.json", f"{R}_clocks_hopf3_f64.json"),
    (f"{R}_clocks_ring_f32.json", f"{R}_clocks_ring_f32.json"), (f"{R}_memcheck.log", f"{R}_memcheck.log"),
]
for src, dst in COPIES:
    if os.path.isfile(os.path.join(G, src)):
        shutil.copy(os.path.join(G, src), os.path.join(P, dst))

# (report, title, algorithmic bytes per launch or None)
REPORTS = [
    ("c2", "C2 create_standard(9) float32, 256 M points — the bench's own launches (80 B/point/direction: 20.48 GB algorithmic per launch)", 256e6 * 80),
    ("hopf3_f64", "C3 create_hopf(3) float64, 128 M points (128 B/point/direction: 16.38 GB)", 128e6 * 128),
    ("stdp8_f64", "C3 create_standard_prime(8) float64, 128 M points (144 B/point/direction: 18.43 GB)", 128e6 * 144),
    ("ring", "C4 create_random(32, default_rng(0)) float32, 64 M points, row-ring kernels (264 B/point/direction: 16.90 GB)", 64e6 * 264),
    ("c1", "C1 create_spherical() float64, 1 M points (48 B/point/direction: 48 MB)", 1e6 * 48),
    ("c5", "C5 integrate pieces: broadcast grid kernel (6.79 GB written), row kernel of the fused functional (1.36 GB written), TMA-streamed reduction (1.36 GB read)", None),
]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active", "launch__registers_per_thread",
        "launch__block_size", "launch__grid_size", "launch__shared_mem_per_block_dynamic",
        "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
TIME = {"ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0, "second": 1.0, "msecond": 1e-3, "usecond": 1e-6, "nsecond": 1e-9}
out = [f"# {R} — `ncu --set full --clock-control none --import-source on` on the hot kernels at the BASELINE sizes", "",
       "Every report: the plain command exited 0 first in the same call (tools/evidence_run.sh). Times under the profiler are cold-cache and",
       "serialised; the numbers to read are traffic (DRAM bytes against the algorithmic bytes), pipe / issue utilisation, occupancy, and the",
       "local-memory instruction counts (`smsp__inst_executed_op_local_*`: the deferred-branch stack of round 1 lived there; now 0 in every loop).", ""]
traffic = {}
for name, what, alg in REPORTS:
    path = os.path.join(G, f"prof_{R}_{name}.ncu-rep")
    if not os.path.isfile(path):
        continue
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    out += [f"## {what}", ""]
    for r in rows[2:]:
        kname = r[hdr.index("Kernel Name")]
        out += [f"### `{kname}`", "", "| metric | value | unit |", "|---|---|---|"]
        for k in KEYS:
            if k in hdr:
                out.append(f"| {k} | {r[hdr.index(k)]} | {units[hdr.index(k)]} |")
        rd = float(r[hdr.index("dram__bytes_read.sum")]) * SCALE[units[hdr.index("dram__bytes_read.sum")]]
        wr = float(r[hdr.index("dram__bytes_write.sum")]) * SCALE[units[hdr.index("dram__bytes_write.sum")]]
        dur = float(r[hdr.index("gpu__time_duration.sum")]) * TIME[units[hdr.index("gpu__time_duration.sum")]]
        line = f"| **DRAM traffic** | {(rd + wr) / 1e9:.3f} GB in {dur * 1e3:.3f} ms = {(rd + wr) / dur / 1e9:.0f} GB/s under the profiler |"
        line += f" algorithmic {alg / 1e9:.3f} GB, ratio {(rd + wr) / alg:.3f} |" if alg else " |"
        out += [line, ""]
        if name == "c2":
            key = "us::from_cartesian" if "from_" in kname else "us::to_cartesian"
            traffic[key] = {"bytes_per_launch": rd + wr, "algorithmic_bytes": alg, "ratio": (rd + wr) / alg, "points": 256_000_000,
                            "source": f"profiles/{R}_ncu.md (ncu --set full of the bench's own launch, 256 M points)"}
    # where the warp samples sit
    stats = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_source_stats.py"), path, "8"], capture_output=True, text=True).stdout
    seen = set()
    keep = []
    for block in stats.split("## ")[1:]:
        head = block.splitlines()[0]
        if head in seen:
            continue
        seen.add(head)
        keep.append("## " + block.rstrip())
    out += ["Source page (warp samples by stall reason, executed opcode mix, hottest instructions):", "", "```", *keep, "```", ""]
open(os.path.join(P, f"{R}_ncu.md"), "w").write("\n".join(out) + "\n")
if traffic:
    json.dump(traffic, open(os.path.join(P, f"traffic_{R}.json"), "w"), indent=1)

# launch list
lp = os.path.join(P, f"{R}_launches.csv")
if os.path.isfile(lp):
    rows = list(csv.reader(l for l in open(lp) if not l.startswith("==")))
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot, cnt = collections.OrderedDict(), collections.Counter()

    def ms(r):
        v = float(r[vi].replace(",", ""))
        return v * TIME.get(r[ui], 1e-9) * 1e3

    for r in rows[1:]:
        if len(r) > vi:
            tot[r[ki]] = tot.get(r[ki], 0) + ms(r)
            cnt[r[ki]] += 1
    all_ms = sum(tot.values())
    L = [f"# {R} — ncu launch list of `python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-quadrature --no-configs --no-sustained`", "",
         "`ncu --metrics gpu__time_duration.sum --clock-control none -c 80` (cold-cache, serialised: compare SHARES).", "",
         "| kernel | launches | total ms | share |", "|---|---|---|---|"]
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        L.append(f"| `{k[:100]}` | {cnt[k]} | {v:.3f} | {v / all_ms:.1%} |")
    f = [ms(r) for r in rows[1:] if len(r) > vi and "from_cartesian_tiled" in r[ki]]
    t = [ms(r) for r in rows[1:] if len(r) > vi and "to_cartesian_tiled" in r[ki]]
    if f and t and os.path.isfile(os.path.join(P, f"{R}_bench.json")):
        rf = json.load(open(os.path.join(P, f"{R}_bench.json")))["roofline"]["other_kernel"]
        fm, tm = sum(f) / len(f), sum(t) / len(t)
        own = sum(v for k, v in tot.items() if "us::" in k)
        L += ["", f"Library kernels (`us::*`): {own:.3f} ms of {all_ms:.3f} ms; the rest is torch input generation (`rand`, `mul_`, `sub_`) and the bench's own "
              "round-trip check, all outside the timed region. Inside the timed region only the two library kernels launch (2 per step).",
              f"Per launch under ncu: from_cartesian_tiled {fm:.3f} ms ({fm / (fm + tm):.1%} of a step), to_cartesian_tiled {tm:.3f} ms ({tm / (fm + tm):.1%}); "
              f"bench.py (CUDA events, no profiler, mean of 20 steps): {rf['from_cartesian_ms']:.3f} ms ({rf['from_cartesian_ms'] / (rf['from_cartesian_ms'] + rf['to_cartesian_ms']):.1%}) / "
              f"{rf['to_cartesian_ms']:.3f} ms ({rf['to_cartesian_ms'] / (rf['from_cartesian_ms'] + rf['to_cartesian_ms']):.1%})."]
    open(os.path.join(P, f"{R}_launches.md"), "w").write("\n".join(L) + "\n")

# SASS evidence and ptxas log of the shipped build
sass = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sass_excerpt.py")], capture_output=True, text=True).stdout
open(os.path.join(P, f"{R}_sass_hot.txt"), "w").write(sass)
shutil.copy(os.path.join(ROOT, "ultrasphere_b200", "lib", "ptxas.log"), os.path.join(P, f"{R}_ptxas.log"))
print("wrote", sorted(f for f in os.listdir(P) if f.startswith(R) or f.startswith("traffic_" + R)))
