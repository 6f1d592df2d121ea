"""``random_ball``: uniform points in the unit ball / on the unit sphere, generated ON THE DEVICE.

Same signature and return types as the reference (/root/reference/src/ultrasphere/_random.py:96-192),
which samples with NumPy on the host and copies ``c_ndim * N`` values to the device — at 10⁸ points
that copy would dominate any pipeline that feeds the transforms.  Here a counter-based Philox
generator runs in the kernel (``us_random_ball`` / ``us_random_uniform``).

The *distribution* is the reference's (Gaussian / exponential construction of Barthe et al. for
``type="uniform"``, per-angle uniform ranges for ``type="spherical"``); the *stream* is not NumPy's
PCG64, so values differ from the reference for the same ``rng``.  Determinism is kept: the seed is
drawn from ``rng``, so the same generator state gives the same points.
"""
from __future__ import annotations

import ctypes as C
import math
from collections.abc import Sequence
from typing import Any, Literal

import numpy as np
import torch

from . import _native
from ._coordinates import BranchingType, SphericalCoordinates
from ._integral import _check_xp
from ._transform import _DTYPE_CODE, _stream_ptr, device_guard

_RANGES = {
    BranchingType.A: (0.0, 2 * math.pi),
    BranchingType.B: (0.0, math.pi),
    BranchingType.BP: (-math.pi / 2, math.pi / 2),
    BranchingType.C: (0.0, math.pi / 2),
}


def _resolve(device: Any, dtype: Any) -> tuple[torch.device, torch.dtype]:
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError(f"device={dev}: ultrasphere_b200 samples on CUDA devices only (no CPU fallback)")
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    dt = torch.float64 if dtype is None else dtype  # the reference hands float64 NumPy arrays to asarray
    if dt not in _DTYPE_CODE:
        raise TypeError(f"unsupported dtype {dt}: float32 or float64")
    return dev, dt


def random_ball(
    c: SphericalCoordinates[Any, Any],
    *,
    shape: Sequence[int],
    xp: Any,
    device: Any | None = None,
    dtype: Any | None = None,
    type: Literal["uniform", "spherical"] = "uniform",  # noqa: A002
    rng: np.random.Generator | None = None,
    surface: bool = False,
) -> Any:
    """Random points in the unit ball (``surface=False``) or on the unit sphere.

    ``type="uniform"``: a ``(c_ndim, *shape)`` tensor of points uniformly distributed in the ball /
    on the sphere.  ``type="spherical"``: a dict with every angle uniform on its range
    (a: [0, 2π), b: [0, π), b': [-π/2, π/2), c: [0, π/2)) and, unless ``surface``, ``"r"`` uniform on [0, 1).
    """
    _check_xp(xp)
    dev, dt = _resolve(device, dtype)
    rng = np.random.default_rng() if rng is None else rng
    seed = int(rng.integers(0, 2**63 - 1))
    shape = tuple(int(e) for e in shape)
    n = 1
    for e in shape:
        n *= e
    lib = _native.lib()
    if type == "uniform":
        out = torch.empty((c.c_ndim, *shape), dtype=dt, device=dev)
        if n:
            row = n * out.element_size()
            ptrs = _native.void_p_array([out.data_ptr() + i * row for i in range(c.c_ndim)])
            with device_guard(dev):
                status = lib.us_random_ball(_DTYPE_CODE[dt], c.c_ndim, n, seed, 1 if surface else 0, ptrs, _stream_ptr(dev))
            _native.check(status, "us_random_ball")
        return out
    if type == "spherical":
        keys: list[Any] = list(c.s_nodes)
        lo = [_RANGES[c.branching_types[k]][0] for k in keys]
        hi = [_RANGES[c.branching_types[k]][1] for k in keys]
        if not surface:
            keys.append("r")
            lo.append(0.0)
            hi.append(1.0)
        if not keys:
            return {}
        out = torch.empty((len(keys), *shape), dtype=dt, device=dev)
        if n:
            row = n * out.element_size()
            ptrs = _native.void_p_array([out.data_ptr() + i * row for i in range(len(keys))])
            with device_guard(dev):
                status = lib.us_random_uniform(
                    _DTYPE_CODE[dt], len(keys), n, seed, (C.c_double * len(keys))(*lo), (C.c_double * len(keys))(*hi), ptrs, _stream_ptr(dev)
                )
            _native.check(status, "us_random_uniform")
        return dict(zip(keys, out.unbind(0)))
    raise ValueError(f"Invalid type {type}.")
