"""torch custom-op layer over the C ABI.

``torch.ops.ultrasphere_b200.{to_cartesian, from_cartesian, weighted_reduce}`` are thin
``torch.library`` registrations whose CUDA implementations are the very same launches the
``SphericalCoordinates`` methods make (``_transform._launch_*`` → ``us_to_cartesian`` /
``us_from_cartesian`` / ``us_weighted_reduce``).  They exist so that the path composes with the
rest of the PyTorch stack — ``torch.compile`` / FX tracing see one opaque op with a shape-inferring
fake implementation instead of Python glue, and dispatcher tooling (profiler, ``opcheck``) sees named
ops.  Custom-op schemas cannot carry dicts or Python objects, so a coordinate tree travels as an
integer handle into a registry of compiled plans, and dict inputs/outputs are flattened to
``Tensor[]`` in ``s_nodes`` / ``c_nodes`` order by the caller.

There is no CPU implementation on purpose: calling the ops with CPU tensors raises
``NotImplementedError`` from the dispatcher.
"""
from __future__ import annotations

import weakref
from typing import Any

import torch

from ._coordinates import SphericalCoordinates
from ._plan import CompiledTree, plan_for
from ._transform import broadcast_shapes

_plans: dict[int, CompiledTree] = {}
_handles: "weakref.WeakKeyDictionary[Any, int]" = weakref.WeakKeyDictionary()


def plan_handle(c: SphericalCoordinates[Any, Any]) -> int:
    """Integer handle of ``c``'s compiled tree (stable for the life of the object)."""
    h = _handles.get(c)
    if h is None:
        h = len(_plans) + 1
        _plans[h] = plan_for(c)
        _handles[c] = h
    return h


def _plan(handle: int) -> CompiledTree:
    try:
        return _plans[handle]
    except KeyError:
        raise ValueError(f"unknown ultrasphere_b200 plan handle {handle}") from None


@torch.library.custom_op("ultrasphere_b200::to_cartesian", mutates_args=(), device_types="cuda")
def to_cartesian_op(angles: list[torch.Tensor], r: torch.Tensor | None, plan: int) -> torch.Tensor:
    """angles: s_nodes order; returns the stacked ``(c_ndim, *broadcast_shape)`` Cartesian array."""
    from ._transform import _compute_dtype, _launch_to

    ct = _plan(plan)
    if len(angles) != ct.n_nodes:
        raise ValueError(f"expected {ct.n_nodes} angle tensors, got {len(angles)}")
    ops = list(angles) + ([r] if r is not None else [])
    dtype = _compute_dtype(ops)
    ang = [a if a.dtype == dtype else a.to(dtype) for a in angles]
    rr = None if r is None else (r if r.dtype == dtype else r.to(dtype))
    full = tuple(broadcast_shapes(*[t.shape for t in ops]))
    return _launch_to(ct, angles[0].device, dtype, ang, rr, full, list(range(ct.n_leaves)))


@to_cartesian_op.register_fake
def _(angles: list[torch.Tensor], r: torch.Tensor | None, plan: int) -> torch.Tensor:
    ct = _plan(plan)
    ops = list(angles) + ([r] if r is not None else [])
    dtype = ops[0].dtype
    for t in ops[1:]:
        dtype = torch.promote_types(dtype, t.dtype)
    full = broadcast_shapes(*[t.shape for t in ops])
    return angles[0].new_empty((ct.n_leaves, *full), dtype=dtype)


@torch.library.custom_op("ultrasphere_b200::from_cartesian", mutates_args=(), device_types="cuda")
def from_cartesian_op(leaves: list[torch.Tensor], plan: int) -> torch.Tensor:
    """leaves: c_nodes order; returns ``(1 + s_ndim, *broadcast_shape)``: r, then angles in s_nodes order."""
    from ._transform import _compute_dtype, _launch_from

    ct = _plan(plan)
    if len(leaves) != ct.n_leaves:
        raise ValueError(f"expected {ct.n_leaves} leaf tensors, got {len(leaves)}")
    dtype = _compute_dtype(leaves)
    xs = [x if x.dtype == dtype else x.to(dtype) for x in leaves]
    full = tuple(broadcast_shapes(*[t.shape for t in xs]))
    return _launch_from(ct, leaves[0].device, dtype, xs, full, True, list(range(ct.n_nodes)))


@from_cartesian_op.register_fake
def _(leaves: list[torch.Tensor], plan: int) -> torch.Tensor:
    ct = _plan(plan)
    dtype = leaves[0].dtype
    for t in leaves[1:]:
        dtype = torch.promote_types(dtype, t.dtype)
    full = broadcast_shapes(*[t.shape for t in leaves])
    return leaves[0].new_empty((ct.n_nodes + 1, *full), dtype=dtype)


@torch.library.custom_op("ultrasphere_b200::weighted_reduce", mutates_args=(), device_types="cuda")
def weighted_reduce_op(val: torch.Tensor, weights: list[torch.Tensor]) -> torch.Tensor:
    """sum over the leading len(weights) axes of prod_i w_i[idx_i] * val."""
    from ._integral import weighted_reduce

    return weighted_reduce(val, list(weights))


@weighted_reduce_op.register_fake
def _(val: torch.Tensor, weights: list[torch.Tensor]) -> torch.Tensor:
    dtype = torch.promote_types(val.dtype, weights[0].dtype)
    return val.new_empty(val.shape[len(weights) :], dtype=dtype)
