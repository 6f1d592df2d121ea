"""The N>1 host logic on the CPU: two gloo ranks shard the leading quadrature axis exactly as
``integrate(..., group=pg)`` does (``shard_rows``), compute their partial integrals with the NumPy
oracle, and combine them with the product's ``_finish`` (one all-reduce).  The CUDA kernels are not
involved; what is checked is the partitioning, the collective and that the shares add up to the
single-process integral."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import vks_oracle as vo
from ultrasphere_b200._integral import _finish, shard_rows


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, spec: str, n: int, out) -> None:
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        t = vo.make(spec)
        k = np.random.default_rng(0).uniform(0, 1, t.c_ndim)
        xs, ws = vo.roots(t, n, expand_dims_x=True)
        first = t.s_nodes[0]
        lo, hi = shard_rows(ws[first].shape[0], rank, world)
        part_x = dict(xs)
        part_x[first] = xs[first][lo:hi]
        part_w = dict(ws)
        part_w[first] = ws[first][lo:hi]
        if hi > lo:
            val = np.exp(np.einsum("v,v...->...", k, vo.to_cartesian(t, part_x, as_array=True)))
            part = vo.contract(t, val, part_w)
        else:
            part = np.zeros(())
        total = _finish(torch.as_tensor(np.asarray(part, dtype=np.float64)), dist.group.WORLD, sharded=True)
        out[rank] = float(total)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("spec,n,world", [("standard:4", 8, 2), ("hopf:2", 6, 2), ("spherical", 1, 3)])
def test_sharded_integral_adds_up(spec, n, world):
    t = vo.make(spec)
    k = np.random.default_rng(0).uniform(0, 1, t.c_ndim)
    want = float(vo.integrate(t, lambda s: np.exp(np.einsum("v,v...->...", k, vo.to_cartesian(t, s, as_array=True))), False, n))
    manager = mp.Manager()
    out = manager.dict()
    mp.spawn(_worker, args=(world, _free_port(), spec, n, out), nprocs=world, join=True)
    assert len(out) == world
    for rank in range(world):
        assert abs(out[rank] - want) <= 1e-13 * abs(want), (rank, out[rank], want)


def test_shard_rows_partition():
    for n in (1, 2, 7, 96, 192):
        for world in (1, 2, 3, 4, 8):
            parts = [shard_rows(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
            sizes = [hi - lo for lo, hi in parts]
            assert max(sizes) - min(sizes) <= 1
