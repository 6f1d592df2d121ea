"""BASELINE.json configs[4]: us.integrate on create_standard(4), n=96 (169,869,312 nodes), float64,
f = exp(k.x), leading grid axis sharded over the ranks, one NCCL all-reduce.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/integrate_multi.py
    python tools/integrate_multi.py            # single GPU

Prints one JSON line from rank 0: value, analytic value, relative error, end-to-end nodes/s.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ultrasphere_b200 as us  # noqa: E402


def main() -> None:
    os.environ.setdefault("NCCL_DEBUG", "WARN")  # no version banner on stdout
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD
    n = int(os.environ.get("US_N", "96"))
    c = us.create_standard(4)
    k5 = np.random.default_rng(0).uniform(0, 1, 5)
    kd = torch.from_numpy(k5).to(dev)

    fused = os.environ.get("US_FUSED", "0") == "1"  # k.x through to_cartesian_dot: the Cartesian grid is never written

    def f(s):
        if fused:
            return torch.exp(c.to_cartesian_dot(s, kd))
        return torch.exp(torch.einsum("v,v...->...", kd, c.to_cartesian(s, as_array=True)))

    def run():
        return us.integrate(c, f, False, n, xp=torch, device=dev, dtype=torch.float64, group=group)

    for _ in range(3):
        got = run()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = 10
    e0.record()
    for _ in range(steps):
        got = run()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    from scipy.special import iv

    kn = float(np.linalg.norm(k5))
    want = float((2 * np.pi) ** 2.5 * kn ** (1 - 2.5) * iv(1.5, kn))  # closed form of the integral of exp(k.x) over S^4
    nodes = n * n * n * 2 * n
    if rank == 0:
        print(json.dumps({
            "workload": f"integrate(create_standard(4), exp(k.x), n={n}) float64", "integrand": "exp(to_cartesian_dot(s, k))" if fused else "exp(einsum(k, to_cartesian(s, as_array=True)))", "n_gpus": world, "nodes": nodes,
            "value": got.item(), "analytic": want, "rel_err": abs(got.item() - want) / want,
            "ms_per_integral": float(ms.item()), "nodes_per_s": nodes / (float(ms.item()) * 1e-3),
        }))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
