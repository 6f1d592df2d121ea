"""A/B builds of the library on ONE box (boxes of the pool differ by a few percent).

    python tools/ab.py --libs ab_libs/r01.so ab_libs/new.so --cases standard:9:f32:256000000 hopf:3:f64:128000000 [--env US_TILE_GROUPS=5]

Every (library, case) pair runs in its own process (ULTRASPHERE_B200_LIB selects the build), the
libraries alternate, two rounds; prints one line per run and a table of the best times with the
fraction of the measured HBM peak."""
import argparse, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r"""
import json, os, sys, torch
sys.path.insert(0, %r)
import ultrasphere_b200 as us
from tests.helpers import make
spec, dt, n = sys.argv[1], sys.argv[2], int(sys.argv[3])
tdt = torch.float32 if dt == "f32" else torch.float64
c = make(us, spec)
x = torch.rand((c.c_ndim, n), device="cuda", dtype=tdt) * 2 - 1
s = c.from_cartesian(x)
def t(fn, reps=7):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    best = 1e9; tot = 0.0
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        ms = e0.elapsed_time(e1); best = min(best, ms); tot += ms
    return best, tot / reps
fb, fm = t(lambda: c.from_cartesian(x))
tb, tm = t(lambda: c.to_cartesian(s, as_array=True))
y = c.to_cartesian(s, as_array=True)
err = float((y - x).abs().max())
print(json.dumps({"from_best": fb, "from_mean": fm, "to_best": tb, "to_mean": tm, "err": err, "bpp": 2 * c.c_ndim * x.element_size()}))
""" % ROOT


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--libs", nargs="+", required=True)
    ap.add_argument("--cases", nargs="+", required=True, help="spec:dtype:n with spec itself possibly containing ':' (standard:9:f32:1000000)")
    ap.add_argument("--env", nargs="*", default=[], help="NAME=VALUE applied to every run; NAME=a,b sweeps values as extra variants")
    ap.add_argument("--rounds", type=int, default=2)
    args = ap.parse_args()
    peak = 6536.4
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(pk):
        peak = json.load(open(pk))["hbm_gbs"]
    variants = []
    sweeps = [(e.split("=", 1)[0], e.split("=", 1)[1].split(",")) for e in args.env]
    for lib in args.libs:
        combos = [{}]
        for name, values in sweeps:
            combos = [dict(c, **{name: v}) for c in combos for v in values]
        for combo in combos:
            label = os.path.basename(lib) + "".join(f" {k}={v}" for k, v in combo.items())
            variants.append((label, lib, combo))
    best: dict = {}
    for rnd in range(args.rounds):
        for case in args.cases:
            *spec, dt, n = case.split(":")
            spec = ":".join(spec)
            for label, lib, combo in variants:
                env = dict(os.environ, ULTRASPHERE_B200_LIB=os.path.abspath(lib), **combo)
                r = subprocess.run([sys.executable, "-c", CHILD, spec, dt, n], env=env, capture_output=True, text=True, timeout=600)
                if r.returncode != 0:
                    print(f"FAILED {label} {case}: {r.stderr[-400:]}", flush=True)
                    continue
                res = json.loads(r.stdout.strip().splitlines()[-1])
                print(f"{case:34s} {label:40s} from {res['from_best']:.4f} ({res['from_mean']:.4f})  to {res['to_best']:.4f} ({res['to_mean']:.4f})  err {res['err']:.2e}", flush=True)
                key = (case, label)
                old = best.get(key)
                if old is None:
                    best[key] = res
                else:
                    for k in ("from_best", "to_best"):
                        old[k] = min(old[k], res[k])
    print("\n| case | build | from ms | from frac | to ms | to frac |\n|---|---|---|---|---|---|")
    for (case, label), res in best.items():
        n = int(case.rsplit(":", 1)[1])
        gb = n * res["bpp"] / 1e9
        print(f"| {case} | {label} | {res['from_best']:.4f} | {gb / res['from_best'] * 1e3 / peak:.3f} | {res['to_best']:.4f} | {gb / res['to_best'] * 1e3 / peak:.3f} |")


if __name__ == "__main__":
    main()
