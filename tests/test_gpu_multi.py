"""integrate(..., group=pg) on real GPUs: two NCCL ranks shard the leading quadrature axis and
all-reduce their partial integrals.  Skipped on a box with fewer than two GPUs (the host-side
partitioning logic is covered on the CPU by tests/test_multi_gpu_cpu.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import vks_oracle as vo

pytestmark = pytest.mark.gpu


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, out) -> None:
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), NCCL_DEBUG="WARN")
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        import ultrasphere_b200 as us

        res = {}
        for spec, c, n in (("standard:4", us.create_standard(4), 24), ("hopf:2", us.create_hopf(2), 17)):
            k = torch.from_numpy(np.random.default_rng(0).uniform(0, 1, c.c_ndim)).to(dev)
            plain = lambda s: torch.exp(torch.einsum("v,v...->...", k, c.to_cartesian(s, as_array=True)))  # noqa: E731
            fused = lambda s: torch.exp(c.to_cartesian_dot(s, k))  # noqa: E731
            kw = dict(xp=torch, device=dev, dtype=torch.float64, group=dist.group.WORLD)
            res[spec] = (us.integrate(c, plain, False, n, **kw).item(), us.integrate(c, fused, False, n, slab_bytes=1 << 16, **kw).item())
        # batch sharding of the transforms needs no collective: each rank its own slice
        c = us.create_standard(9)
        x = torch.rand((10, 100_000), device=dev, generator=torch.Generator(device=dev).manual_seed(7)) * 2 - 1
        lo, hi = rank * 50_000, (rank + 1) * 50_000
        back = c.to_cartesian(c.from_cartesian(x[:, lo:hi].contiguous()), as_array=True)
        res["round_trip_err"] = (back - x[:, lo:hi]).abs().max().item()
        out[rank] = res
    finally:
        dist.destroy_process_group()


def test_two_rank_sharded_integrals_and_batches():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    manager = mp.Manager()
    out = manager.dict()
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    assert len(out) == 2
    for spec, n in (("standard:4", 24), ("hopf:2", 17)):
        t = vo.make(spec)
        k = np.random.default_rng(0).uniform(0, 1, t.c_ndim)
        want = float(vo.integrate(t, lambda s: np.exp(np.einsum("v,v...->...", k, vo.to_cartesian(t, s, as_array=True))), False, n))
        for rank in range(2):
            for got in out[rank][spec]:
                assert abs(got - want) <= 1e-12 * abs(want), (spec, rank, got, want)
    assert max(out[r]["round_trip_err"] for r in range(2)) <= 2e-6
