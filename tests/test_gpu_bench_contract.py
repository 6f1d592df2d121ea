"""bench.py on a real GPU: ONE JSON line with every key of the contract (a reduced batch keeps it short)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu


def test_bench_prints_one_contract_line():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    done = subprocess.run(
        [sys.executable, os.path.join(root, "bench.py"), "--steps", "3", "--warmup", "3", "--points", "8000000", "--no-cpu-baseline", "--leg-scale", "0.05", "--sustain-seconds", "0.1"],
        capture_output=True, text=True, cwd=root, timeout=900,
    )
    assert done.returncode == 0, done.stderr[-3000:]
    lines = [l for l in done.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    line = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "roofline", "e2e", "gpu_launches", "clocks", "quadrature", "configs"):
        assert key in line, key
    assert line["n_gpus"] == 1 and line["steps"] == 3 and line["gpu_launches"] == 6 and line["dtype"] == "f32"
    roof = line["roofline"]
    assert roof["bound"] == "hbm" and roof["unit"] == "GB/s" and 0 < roof["frac"] < 1.05
    assert abs(roof["frac"] - roof["achieved"] / roof["peak"]) < 1e-9
    e2e = line["e2e"]
    assert e2e["value"] > 0 and e2e["h2d_bytes_per_step"] > 0 and e2e["d2h_bytes_per_step"] > 0 and e2e["value"] < line["value"]
    assert line["clocks"]["sm_mhz"] > 0
    quad = line["quadrature"]
    assert quad["reference_form"]["rel_err_vs_analytic"] <= 1e-12 and quad["fused_functional"]["rel_err_vs_analytic"] <= 1e-12
    assert quad["fused_functional"]["nodes_per_s"] > quad["reference_form"]["nodes_per_s"] > 0
    assert quad["fused_functional_graph"]["rel_err_vs_analytic"] <= 1e-12
    assert roof["copy_in_this_run"]["sustained_copy_gbs"] > 0 and roof["traffic"] > 0
    assert e2e["host_link_ceiling"]["d2h_gbs"] > 0 and 0 < e2e["frac_of_link_ceiling"] < 1.2
    legs = line["configs"]
    for key in ("C1", "C3_hopf3", "C3_standard_prime8", "C4"):
        leg = legs[key]
        assert leg["from_cartesian_ms"] > 0 and leg["to_cartesian_ms"] > 0 and 0 < leg["roofline"]["from_cartesian"]["frac"] < 1.1, key
    assert legs["C3_hopf3"]["fp64_pipe"]["from_cartesian"]["frac"] > 0 and "graph_replay" in legs["C1"]
    # the power-settled timings next to the burst ones: headline and every full-size leg
    for rec in (roof["sustained"], legs["C3_hopf3"]["sustained"], legs["C4"]["sustained"]):
        assert rec["from_cartesian"]["ms"] > 0 and 0 < rec["to_cartesian"]["frac"] < 1.1 and rec["value"] > 0
    assert roof["sustained"]["from_cartesian"]["clocks"]["sm_mhz"] > 0
