"""Turn gpurun_out/{launches_rNN.csv, prof_rNN.ncu-rep, bench_rNN*.json} into the committed
summaries under profiles/ (run in the build container, needs the ncu CLI to read the report)."""
import collections, csv, json, os, shutil, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
R = sys.argv[1] if len(sys.argv) > 1 else "r01"
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)
for src, dst in ((f"launches_{R}.csv", f"{R}_launches.csv"), (f"bench_{R}.json", f"{R}_bench.json"), (f"bench_{R}_reference.json", f"{R}_bench_reference.json")):
    if os.path.isfile(os.path.join(G, src)):
        shutil.copy(os.path.join(G, src), os.path.join(P, dst))
raw = subprocess.run(["ncu", "-i", os.path.join(G, f"prof_{R}.ncu-rep"), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
open(os.path.join(P, f"{R}_ncu_full_raw.csv"), "w").write(raw)
bench = json.load(open(os.path.join(P, f"{R}_bench.json")))
rf = bench["roofline"]["other_kernel"]

rows = list(csv.reader(l for l in open(os.path.join(P, f"{R}_launches.csv")) if not l.startswith("==")))
hdr = rows[0]; ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, cnt = collections.OrderedDict(), collections.Counter()
def ms(r):
    v = float(r[vi].replace(",", "")); u = r[ui]
    return v / 1e6 if u == "ns" else v / 1e3 if u == "us" else v * 1e3 if u in ("s", "second") else v
for r in rows[1:]:
    if len(r) > vi:
        tot[r[ki]] = tot.get(r[ki], 0) + ms(r); cnt[r[ki]] += 1
all_ms = sum(tot.values())
L = [f"# {R} — ncu launch list of `python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-quadrature`", "",
     "`ncu --metrics gpu__time_duration.sum --clock-control none -c 80` (cold-cache, serialised: compare SHARES).", "",
     "| kernel | launches | total ms | share |", "|---|---|---|---|"]
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    L.append(f"| `{k[:90]}` | {cnt[k]} | {v:.3f} | {v / all_ms:.1%} |")
f = [ms(r) for r in rows[1:] if len(r) > vi and "from_cartesian_tiled" in r[ki]]
t = [ms(r) for r in rows[1:] if len(r) > vi and "to_cartesian_tiled" in r[ki]]
fm, tm = sum(f) / len(f), sum(t) / len(t)
own = sum(v for k, v in tot.items() if "cartesian" in k)
L += ["", f"Library kernels (`us::*cartesian*`): {own:.3f} ms of {all_ms:.3f} ms; the rest is torch input generation (`rand`, `mul_`, `sub_`) and the bench's own round-trip check, all outside the timed region.",
      f"Inside the timed region only the two library kernels launch (2 per step). Per launch under ncu: from_cartesian_tiled {fm:.3f} ms ({fm / (fm + tm):.1%} of a step), to_cartesian_tiled {tm:.3f} ms ({tm / (fm + tm):.1%}); "
      f"bench.py (CUDA events, no profiler): {rf['from_cartesian_ms']:.3f} ms ({rf['from_cartesian_ms'] / (rf['from_cartesian_ms'] + rf['to_cartesian_ms']):.1%}) / {rf['to_cartesian_ms']:.3f} ms ({rf['to_cartesian_ms'] / (rf['from_cartesian_ms'] + rf['to_cartesian_ms']):.1%})."]
open(os.path.join(P, f"{R}_launches.md"), "w").write("\n".join(L) + "\n")

rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"]
out = [f"# {R} — `ncu --set full --clock-control none` on the two tiled kernels", "",
       "Command: `ncu --set full --clock-control none --import-source on -k regex:cartesian_tiled -s 4 -c 2 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-quadrature --points 64000000` "
       "(create_standard(9), float32, 64 M points; algorithmic bytes per launch = 64e6 x 80 B = 5.12 GB). Raw page: `" + f"{R}_ncu_full_raw.csv`.", ""]
traffic = {}
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    out += [f"## `{name}`", "", "| metric | value | unit |", "|---|---|---|"]
    for k in keys:
        if k in hdr:
            out.append(f"| {k} | {r[hdr.index(k)]} | {units[hdr.index(k)]} |")
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}
    tr = float(r[hdr.index("dram__bytes_read.sum")]) * scale[units[hdr.index("dram__bytes_read.sum")]] + float(r[hdr.index("dram__bytes_write.sum")]) * scale[units[hdr.index("dram__bytes_write.sum")]]
    inst = float(r[hdr.index("smsp__inst_executed.sum")])
    out += [f"| **traffic (read+write)** | {tr / 1e9:.3f} | GB (algorithmic 5.120 GB, ratio {tr / 5.12e9:.3f}) |",
            f"| **lane-instructions per point** | {inst * 32 / 64e6:.0f} | |", ""]
    traffic["us::from_cartesian" if "from_" in name else "us::to_cartesian"] = {"bytes_per_launch_at_64M_points": tr, "algorithmic_bytes": 5.12e9, "ratio": tr / 5.12e9}
open(os.path.join(P, f"{R}_ncu_full.md"), "w").write("\n".join(out) + "\n")
json.dump(traffic, open(os.path.join(P, f"traffic_{R}.json"), "w"), indent=1)
shutil.copy(os.path.join(ROOT, "ultrasphere_b200", "lib", "ptxas.log"), os.path.join(P, f"{R}_ptxas.log"))
print("\n".join(L[-3:])); print(json.dumps(traffic))
