"""Tensor-product Gauss–Jacobi quadrature on the sphere: ``roots`` and ``integrate``.

Same signatures and results as /root/reference/src/ultrasphere/_integral.py:18-116 (``roots``)
and :157-276 (``integrate``) with ``xp=torch`` on a CUDA device.

* The 1-D rules come from the very same host call the reference makes
  (``scipy.special.roots_jacobi``; trapezoid for type ``a``) so nodes and weights are bit-identical;
  the reference's argument order for type ``c`` — ``(S_cos/2, S_sin/2)`` as ``(alpha, beta)`` — is
  reproduced on purpose.  ``rule_backend("device")`` builds all tables of a
  coordinate system in one kernel launch instead (``us_quadrature_rules``: no scipy call, no H2D
  copy; nodes within 1 ulp of scipy's, weights more accurate than scipy's) — for workloads of
  many small integrals whose ``n`` changes from call to call.
* The contraction ``sum_grid prod_i w_i * f`` is ONE fused pass over the integrand's values
  (``us_weighted_reduce``: warp-shuffle + block tree, float64 accumulation, fixed cross-block
  order) instead of ``s_ndim`` successive ``vecdot`` passes; the separable branch is one launch of
  ``us_separable_reduce`` for all axes.
* ``f`` is evaluated in slabs of the leading grid axis so that its temporaries stay small, and
  the leading axis can be sharded over the ranks of a ``torch.distributed`` group, the partial
  integrals being combined by a single all-reduce.
"""
from __future__ import annotations

import contextlib
import ctypes as C
import functools
import os
from collections.abc import Callable, Mapping
from typing import Any

import numpy as np
import torch
from scipy.special import roots_jacobi

from . import _native
from ._coordinates import BranchingType, SphericalCoordinates, get_child
from ._transform import _DTYPE_CODE, _stream_ptr, device_guard

# The integrand is evaluated in slabs of the leading grid axis only to bound memory: a slab of
# values is kept below this size unless the caller says otherwise (the c_ndim-times larger Cartesian
# grid most integrands build comes on top).  Measured on B200 (tools/c5_breakdown.py): smaller,
# L2-sized slabs are slower than one pass because of the extra launches, so the default is large.
DEFAULT_SLAB_BYTES = 2 << 30


_RULE_BACKENDS = ("scipy", "device")
_rule_backend = os.environ.get("ULTRASPHERE_B200_RULES", "scipy")
if _rule_backend not in _RULE_BACKENDS:
    raise ValueError(f"ULTRASPHERE_B200_RULES={_rule_backend!r}: expected one of {_RULE_BACKENDS}")


def set_rule_backend(name: str) -> str:
    """Choose where ``roots`` builds the Gauss–Jacobi tables: ``"scipy"`` (host, bit-identical to
    the reference's, cached per ``(n, alpha, beta)``) or ``"device"`` (one kernel launch per
    coordinate system).  Returns the previous setting."""
    global _rule_backend
    if name not in _RULE_BACKENDS:
        raise ValueError(f"rule backend {name!r}: expected one of {_RULE_BACKENDS}")
    previous, _rule_backend = _rule_backend, name
    return previous


@contextlib.contextmanager
def rule_backend(name: str) -> Any:
    """``with rule_backend("device"): ...`` — scoped :func:`set_rule_backend`."""
    previous = set_rule_backend(name)
    try:
        yield
    finally:
        set_rule_backend(previous)


def _check_xp(xp: Any) -> None:
    name = getattr(xp, "__name__", "")
    if xp is torch or name in ("torch", "array_api_compat.torch"):
        return
    raise TypeError(
        f"xp={name or xp!r}: ultrasphere_b200 computes on CUDA through torch only (pass xp=torch)"
    )


@functools.lru_cache(maxsize=512)
def _jacobi_rule(kind: str, n: int, alpha: float, beta: float) -> tuple[np.ndarray, np.ndarray]:
    """float64 angles/weights of one Gauss–Jacobi rule, exactly as the reference derives them from
    ``scipy.special.roots_jacobi`` (_integral.py:88-105).  Cached: Golub–Welsch on the host costs
    0.3 ms per angle at n = 96 — a fifth of a whole 170 M-node integral on a B200 otherwise."""
    x, w = roots_jacobi(n, alpha, beta)
    if kind == "b":
        x = np.acos(x)
    elif kind == "b'":
        x = np.asin(x)
    else:
        w = w / 2 ** (alpha + beta + 2)
        x = np.acos(x) / 2
    x.setflags(write=False)
    w.setflags(write=False)
    return x, w


def _rule_parameters(c: SphericalCoordinates[Any, Any], node: Any) -> tuple[str, float, float]:
    """``(kind, alpha, beta)`` of one angle's rule, the Jacobi parameters being what the reference
    hands to ``roots_jacobi`` (_integral.py:88-103)."""
    kind = c.branching_types[node]
    if kind == BranchingType.A:
        return "a", 0.0, 0.0
    if kind == BranchingType.B:
        beta = c.S[get_child(c.G, node, "sin")] / 2
        return "b", beta, beta
    if kind == BranchingType.BP:
        alpha = c.S[get_child(c.G, node, "cos")] / 2
        return "b'", alpha, alpha
    if kind == BranchingType.C:
        # the reference passes (S_cos/2, S_sin/2) as (alpha, beta): reproduced
        return "c", c.S[get_child(c.G, node, "cos")] / 2, c.S[get_child(c.G, node, "sin")] / 2
    raise ValueError(f"Invalid branching type {kind}.")


def _host_rule(c: SphericalCoordinates[Any, Any], node: Any, n: int) -> tuple[np.ndarray, np.ndarray] | None:
    """float64 nodes/weights of one angle on the host; None for type `a` (built on the device)."""
    kind, alpha, beta = _rule_parameters(c, node)
    if kind == "a":
        return None
    return _jacobi_rule(kind, n, alpha, beta)


_KIND_CODE = {"a": 0, "b": 1, "b'": 2, "c": 3}  # US_NODE_* of include/ultrasphere_b200.h


def _rules_on_device(
    c: SphericalCoordinates[Any, Any], n: int, device: Any | None, dtype: Any | None
) -> list[tuple[torch.Tensor, torch.Tensor]]:
    """All 1-D tables of ``c`` from one launch of ``us_quadrature_rules`` (s_nodes order)."""
    if int(n) != n or n < 1:
        raise ValueError("n must be a positive integer.")
    n = int(n)
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    if device.type != "cuda":
        raise RuntimeError(f"device={device}: the device rule backend builds its tables on CUDA devices only")
    native = dtype in (None, torch.float32, torch.float64)
    default = torch.get_default_dtype() if torch.get_default_dtype() in _DTYPE_CODE else torch.float32
    kinds, counts, alphas, betas, dtypes = [], [], [], [], []
    for node in c.s_nodes:
        kind, alpha, beta = _rule_parameters(c, node)
        # dtype=None in the reference: `arange(2n) * pi / n` is torch's default float type, the
        # scipy tables stay float64
        t = (dtype if dtype is not None else (default if kind == "a" else torch.float64)) if native else torch.float64
        count = 2 * n if kind == "a" else n
        if kind != "a" and count > _native.US_MAX_RULE_POINTS:
            raise ValueError(
                f"n={n}: the device rule backend holds at most {_native.US_MAX_RULE_POINTS} Gauss points per angle "
                '(use set_rule_backend("scipy"))'
            )
        kinds.append(_KIND_CODE[kind])
        counts.append(count)
        alphas.append(float(alpha))
        betas.append(float(beta))
        dtypes.append(t)
    # one allocation per dtype; the tables are 16-byte aligned views into it (x then w of each angle)
    steps = [(k + 3) & ~3 for k in counts]
    pools = {t: torch.empty(2 * sum(k for k, u in zip(steps, dtypes) if u == t), dtype=t, device=device) for t in set(dtypes)}
    used = dict.fromkeys(pools, 0)
    tables = []
    for count, step, t in zip(counts, steps, dtypes):
        at = used[t]
        tables.append((pools[t][at : at + count], pools[t][at + step : at + step + count]))
        used[t] = at + 2 * step
    dtypes = [_DTYPE_CODE[t] for t in dtypes]
    lib = _native.lib()
    for lo in range(0, len(kinds), _native.US_MAX_AXES):
        hi = min(len(kinds), lo + _native.US_MAX_AXES)
        k = hi - lo
        with device_guard(device):
            status = lib.us_quadrature_rules(
                k,
                (C.c_uint8 * k)(*kinds[lo:hi]),
                _native.i64_array(counts[lo:hi]),
                (C.c_double * k)(*alphas[lo:hi]),
                (C.c_double * k)(*betas[lo:hi]),
                (C.c_int32 * k)(*dtypes[lo:hi]),
                _native.void_p_array([x.data_ptr() for x, _ in tables[lo:hi]]),
                _native.void_p_array([w.data_ptr() for _, w in tables[lo:hi]]),
                _stream_ptr(device),
            )
        _native.check(status, "us_quadrature_rules")
    if not native:
        tables = [(x.to(dtype), w.to(dtype)) for x, w in tables]
    return tables


def roots(
    c: SphericalCoordinates[Any, Any],
    n: int,
    *,
    expand_dims_x: bool,
    expand_dims_w: bool = False,
    device: Any | None = None,
    dtype: Any | None = None,
    xp: Any,
) -> tuple[dict[Any, torch.Tensor], dict[Any, torch.Tensor]]:
    """Quadrature angles and weights per node of ``c.s_nodes`` (1-D, or broadcastable N-D).

    Every call returns fresh tensors (like the reference, _integral.py:108-109): the cached host
    tables are copied, never aliased.  ``device=None`` means the current CUDA device; on a machine
    without one (the tables alone need no kernel) it means the host, as in the reference."""
    _check_xp(xp)
    if device is None and torch.cuda.is_available():
        device = torch.device("cuda", torch.cuda.current_device())
    xs: dict[Any, torch.Tensor] = {}
    ws: dict[Any, torch.Tensor] = {}
    on_device = _rules_on_device(c, n, device, dtype) if _rule_backend == "device" else None
    for i, node in enumerate(c.s_nodes):
        if on_device is not None:
            x, w = on_device[i]
        elif (rule := _host_rule(c, node, n)) is None:
            x = torch.arange(2 * n, device=device, dtype=dtype) * torch.pi / n
            w = torch.ones(2 * n, device=device, dtype=dtype) * torch.pi / n
        else:
            x = torch.tensor(rule[0], device=device, dtype=dtype)  # a copy: the cache stays private
            w = torch.tensor(rule[1], device=device, dtype=dtype)
        index = (None,) * i + (slice(None),) + (None,) * (c.s_ndim - i - 1)
        xs[node] = x[index] if expand_dims_x else x
        ws[node] = w[index] if expand_dims_w else w
    return xs, ws


# --------------------------------------------------------------------------------------
# reductions
# --------------------------------------------------------------------------------------
_scratch: dict[tuple[int, int], torch.Tensor] = {}


def _scratch_for(device: torch.device, nbytes: int) -> torch.Tensor:
    """Per-(device, stream) scratch for the block partials; grown on demand, reused across calls."""
    key = (device.index if device.index is not None else torch.cuda.current_device(), _stream_ptr(device))
    buf = _scratch.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(nbytes, 1 << 16), dtype=torch.uint8, device=device)
        _scratch[key] = buf
    return buf


def _real_view(val: torch.Tensor) -> tuple[torch.Tensor, bool]:
    if val.is_complex():
        return torch.view_as_real(val.contiguous()), True
    return val, False


def _as_supported(val: torch.Tensor, wdtype: torch.dtype) -> torch.Tensor:
    """Cast values to the float/complex type the contraction promotes to (val ⊗ weights)."""
    target = torch.promote_types(val.dtype, wdtype)
    if target == torch.complex32 or target in (torch.float16, torch.bfloat16):
        target = torch.promote_types(target, torch.float32)
    return val if val.dtype == target else val.to(target)


class _WeightedReduceFn(torch.autograd.Function):
    """Autograd wrapper of the fused contraction: it is linear in ``val`` with real weights, so the
    backward pass is the outer product of the weights times the incoming gradient (plain torch
    broadcasting — gradients are not the hot path).  Lets integrals be differentiated with respect
    to parameters of the integrand."""

    @staticmethod
    def forward(ctx: Any, val: torch.Tensor, *weights: torch.Tensor) -> torch.Tensor:
        ctx.save_for_backward(*weights)
        ctx.val_shape = tuple(val.shape)
        ctx.val_dtype = val.dtype
        return _weighted_reduce(val.detach(), [w.detach() for w in weights], out=None)

    @staticmethod
    def backward(ctx: Any, grad_out: torch.Tensor) -> Any:
        weights = ctx.saved_tensors
        k = len(weights)
        grad = grad_out.reshape((1,) * k + tuple(grad_out.shape))
        for i, w in enumerate(weights):
            shape = [1] * len(ctx.val_shape)
            shape[i] = w.numel()
            grad = grad * w.reshape(shape).to(grad_out.real.dtype if grad_out.is_complex() else grad_out.dtype)
        grad = grad.expand(ctx.val_shape)
        if not ctx.val_dtype.is_complex and grad.is_complex():
            grad = grad.real
        return (grad.to(ctx.val_dtype), *([None] * k))


def weighted_reduce(
    val: torch.Tensor,
    weights: list[torch.Tensor],
    *,
    out: torch.Tensor | None = None,
) -> torch.Tensor:
    """``sum over the leading len(weights) axes of prod_i w_i[idx_i] * val`` in one fused pass.

    ``val``: ``(e_0, …, e_{k-1}, *u)`` on a CUDA device with ``e_i == len(w_i)``; returns shape ``u``.
    With ``out`` given the result is accumulated into it (used for slab-wise evaluation) and ``out``
    is returned — except when ``val`` carries an autograd graph: then the sum is returned as a new
    tensor (``out + result``) so that the graph survives.
    """
    if torch.is_grad_enabled():
        if any(w.requires_grad for w in weights):
            raise RuntimeError("weighted_reduce: gradients with respect to the quadrature weights are not provided")
        if val.requires_grad:
            res = _WeightedReduceFn.apply(val, *weights)
            return res if out is None else out + res
    return _weighted_reduce(val, weights, out=out)


def _weighted_reduce(val: torch.Tensor, weights: list[torch.Tensor], *, out: torch.Tensor | None) -> torch.Tensor:
    k = len(weights)
    device = val.device
    if k > _native.US_MAX_DIMS:
        # more grid axes than one launch indexes: contract the leading US_MAX_DIMS axes (the rest
        # rides along as trailing elements), then the remainder the same way
        head = _native.US_MAX_DIMS
        part = _weighted_reduce(val, weights[:head], out=None)
        res = _weighted_reduce(part, weights[head:], out=None)
        if out is None:
            return res
        out += res
        return out
    wdtype = weights[0].dtype
    for w in weights[1:]:
        wdtype = torch.promote_types(wdtype, w.dtype)
    val = _as_supported(val, wdtype)
    rv, is_complex = _real_view(val)
    if not rv.is_contiguous():
        rv = rv.contiguous()
    real_dtype = rv.dtype
    if real_dtype not in _DTYPE_CODE:
        raise TypeError(f"unsupported integrand dtype {val.dtype}")
    ws = [w if (w.dtype == real_dtype and w.is_contiguous()) else w.to(real_dtype).contiguous() for w in weights]
    extent = [int(e) for e in val.shape[:k]]
    for e, w in zip(extent, ws):
        if w.numel() != e:
            raise ValueError(f"weights of length {w.numel()} do not match integrand axis of extent {e}")
    u_shape = tuple(val.shape[k:])
    m = 1
    for e in u_shape:
        m *= e
    m_real = m * (2 if is_complex else 1)
    res_real = torch.empty(u_shape + ((2,) if is_complex else ()), dtype=real_dtype, device=device) if out is None else None
    target = res_real if out is None else (torch.view_as_real(out) if is_complex else out)
    assert target is not None
    if m_real == 0:
        return out if out is not None else (torch.view_as_complex(res_real) if is_complex else res_real)
    total = 1
    for e in extent:
        total *= e
    if total == 0:
        if out is None:
            target.zero_()
        return out if out is not None else (torch.view_as_complex(target) if is_complex else target)
    lib = _native.lib()
    nbytes = int(lib.us_weighted_reduce_scratch_bytes(m_real))
    scratch = _scratch_for(device, nbytes)
    with device_guard(device):
        status = lib.us_weighted_reduce(
            _DTYPE_CODE[real_dtype],
            k,
            _native.i64_array(extent),
            _native.void_p_array([w.data_ptr() for w in ws]),
            rv.data_ptr(),
            m_real,
            target.data_ptr(),
            0 if out is None else 1,
            scratch.data_ptr(),
            scratch.numel(),
            _stream_ptr(device),
        )
    _native.check(status, "us_weighted_reduce")
    if out is not None:
        return out
    return torch.view_as_complex(target) if is_complex else target


def separable_reduce(vals: list[torch.Tensor], weights: list[torch.Tensor]) -> list[torch.Tensor]:
    """``[sum_j w_a[j] * val_a[j, ...] for a]`` in one launch (all axes share a dtype class)."""
    if torch.is_grad_enabled() and any(v.requires_grad for v in vals):
        # the per-axis sum is the one-axis case of the fused contraction, which carries autograd
        return [weighted_reduce(v, [w]) for v, w in zip(vals, weights)]
    device = vals[0].device
    prepared = []
    outs: list[torch.Tensor] = []
    real_dtype = None
    own_dtypes = []
    for v, w in zip(vals, weights):
        v = _as_supported(v, w.dtype)
        own_dtypes.append(v.dtype)  # what the reference's vecdot(w_a, val_a) returns for this axis
        rv, is_complex = _real_view(v)
        rv = rv if rv.is_contiguous() else rv.contiguous()
        prepared.append((rv, is_complex, tuple(v.shape[1:])))
        real_dtype = rv.dtype if real_dtype is None else torch.promote_types(real_dtype, rv.dtype)
    assert real_dtype is not None
    if real_dtype not in _DTYPE_CODE:
        raise TypeError(f"unsupported integrand dtype {real_dtype}")
    lens, ms, wp, vp, op = [], [], [], [], []
    keep = []
    for (rv, is_complex, u_shape), w in zip(prepared, weights):
        if rv.dtype != real_dtype:
            rv = rv.to(real_dtype)
        w = w if (w.dtype == real_dtype and w.is_contiguous()) else w.to(real_dtype).contiguous()
        m_real = 1
        for e in rv.shape[1:]:
            m_real *= e
        o = torch.empty(tuple(rv.shape[1:]), dtype=real_dtype, device=device)
        keep.append((rv, w))
        outs.append(torch.view_as_complex(o) if is_complex else o)
        if m_real == 0:
            continue
        lens.append(int(rv.shape[0]))
        ms.append(m_real)
        wp.append(w.data_ptr())
        vp.append(rv.data_ptr())
        op.append(o.data_ptr())
    lib = _native.lib()
    for lo in range(0, len(lens), _native.US_MAX_AXES):
        hi = min(len(lens), lo + _native.US_MAX_AXES)
        with device_guard(device):
            status = lib.us_separable_reduce(
                _DTYPE_CODE[real_dtype],
                hi - lo,
                _native.i64_array(lens[lo:hi]),
                _native.i64_array(ms[lo:hi]),
                _native.void_p_array(wp[lo:hi]),
                _native.void_p_array(vp[lo:hi]),
                _native.void_p_array(op[lo:hi]),
                _stream_ptr(device),
            )
        _native.check(status, "us_separable_reduce")
    # the launch computes every axis in the widest type present; each result goes back to its own
    return [o if o.dtype == dt else o.to(dt) for o, dt in zip(outs, own_dtypes)]


# --------------------------------------------------------------------------------------
# integrate
# --------------------------------------------------------------------------------------
def _device_rules(
    c: SphericalCoordinates[Any, Any], n: int, device: torch.device, dtype: Any, expand: bool, xp: Any
) -> tuple[dict[Any, torch.Tensor], dict[Any, torch.Tensor]]:
    """Quadrature tables of ``c`` on ``device``, kept on the coordinate object between calls (the
    H2D copies of the tables are synchronous pageable copies, ~20 us each).  The integrand receives
    COPIES of the angle tensors, so an in-place edit cannot poison the cache; the angles of one
    dtype are cached as one flat tensor so that the copy is one launch, not ``s_ndim``."""
    cache = c.__dict__.setdefault("_quad_cache", {})
    key = (n, dtype, device, expand, _rule_backend)
    hit = cache.get(key)
    if hit is None:
        if len(cache) >= 16:
            cache.clear()
        xs, ws = roots(c, n, device=device, dtype=dtype, expand_dims_x=expand, xp=xp)
        pools: dict[torch.dtype, list[Any]] = {}
        layout = []
        for node, x in xs.items():
            members = pools.setdefault(x.dtype, [])
            at = sum(m.numel() for m in members)
            members.append(x.reshape(-1))
            layout.append((node, x.dtype, at, x.numel(), tuple(x.shape)))
        flat = {dt: torch.cat(members) for dt, members in pools.items()}
        hit = (flat, layout, ws)
        cache[key] = hit
    flat, layout, ws = hit
    fresh = {dt: t.clone() for dt, t in flat.items()}
    return {node: fresh[dt][at : at + count].view(shape) for node, dt, at, count, shape in layout}, ws


def _sum_weight(w: torch.Tensor) -> torch.Tensor:
    """1-element weight table for an axis the integrand is constant along (sum of the rule, summed
    in float64, kept in the rule's dtype so that the result dtype follows the reference's promotion)."""
    return w.sum(dtype=torch.float64).to(w.dtype).reshape(1)


def _contract_dense(
    c: SphericalCoordinates[Any, Any],
    val: torch.Tensor,
    ws: list[torch.Tensor],
    out: torch.Tensor | None,
) -> torch.Tensor:
    if val.ndim < c.s_ndim:
        raise ValueError(
            f"The dimension of the return value of f should be at least {c.s_ndim}, got {val.ndim}."
        )
    weights = []
    for i, w in enumerate(ws):
        e = val.shape[i]
        if e == w.numel():
            weights.append(w)
        elif e == 1:
            weights.append(_sum_weight(w))
        else:
            raise ValueError(
                f"integrand axis {i} has extent {e}; expected {w.numel()} (or 1 if f does not depend on it)"
            )
    return weighted_reduce(val, weights, out=out)


def integrate(
    c: SphericalCoordinates[Any, Any],
    f: Callable[[Mapping[Any, torch.Tensor]], Any] | Mapping[Any, torch.Tensor] | torch.Tensor,
    does_f_support_separation_of_variables: bool,
    n: int,
    *,
    xp: Any,
    device: Any | None = None,
    dtype: Any | None = None,
    slab_bytes: int | None = None,
    group: Any | None = None,
) -> Any:
    r"""Integrate ``f`` over the unit sphere :math:`\mathbb{S}^{d-1}` with ``n`` points per angle.

    ``f`` receives the quadrature angles (1-D per node when it supports separation of variables,
    mutually broadcastable ``s_ndim``-D otherwise) and returns values with the grid axes first and
    any extra axes last; or ``f`` is such a value already.  Returns what the reference returns:
    a tensor of the trailing shape, or a dict of per-node 1-D integrals in the separable case.

    Extensions (keyword-only, defaults keep the reference behaviour):
      ``slab_bytes``  evaluate ``f`` on slabs of the leading grid axis of about this many bytes
                      (default 2 GiB; ``0`` = one call)
      ``group``       a ``torch.distributed`` process group: the leading axis is sharded over its
                      ranks and the partial integrals are summed with one all-reduce
    """
    _check_xp(xp)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device())
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError(f"device={device}: ultrasphere_b200 integrates on CUDA devices only (no CPU fallback)")
    xs, ws = _device_rules(c, n, device, dtype, not does_f_support_separation_of_variables, xp)
    nodes = list(c.s_nodes)
    if does_f_support_separation_of_variables or isinstance(f, Mapping):
        val = f(xs) if callable(f) else f
        if isinstance(val, Mapping):
            return _integrate_separable(c, val, ws)
        return _finish(_contract_dense(c, val, [ws[k] for k in nodes], None), group, sharded=False)
    if not callable(f):
        return _finish(_contract_dense(c, f, [ws[k] for k in nodes], None), group, sharded=False)

    # dense, callable: slab the leading axis (and shard it over the group's ranks)
    w_list = [ws[k] for k in nodes]
    n0 = w_list[0].numel()
    lo_rank, hi_rank = 0, n0
    sharded = False
    if group is not None:
        import torch.distributed as dist

        world, rank = dist.get_world_size(group), dist.get_rank(group)
        if world > 1:
            sharded = True
            lo_rank, hi_rank = shard_rows(n0, rank, world)
    rest = 1
    for w in w_list[1:]:
        rest *= w.numel()
    itemsize = w_list[0].element_size()
    budget = DEFAULT_SLAB_BYTES if slab_bytes is None else slab_bytes
    rows = hi_rank - lo_rank
    if budget > 0:
        rows = max(2, min(rows, budget // max(1, rest * itemsize)))
    if rows >= hi_rank - lo_rank and not sharded:
        val = f(xs)
        if isinstance(val, Mapping):
            return _integrate_separable(c, val, ws)
        return _contract_dense(c, val, w_list, None)
    x0 = xs[nodes[0]]
    out: torch.Tensor | None = None
    lo = lo_rank
    while lo < hi_rank:
        hi = min(hi_rank, lo + rows)
        if hi_rank - hi == 1:  # never leave a 1-row slab: it would look like "f ignores this axis"
            hi = hi_rank
        slab = dict(xs)
        slab[nodes[0]] = x0[lo:hi]
        val = f(slab)
        if isinstance(val, Mapping):
            raise TypeError("f returned a mapping but does_f_support_separation_of_variables is False")
        if val.ndim < c.s_ndim:
            raise ValueError(
                f"The dimension of the return value of f should be at least {c.s_ndim}, got {val.ndim}."
            )
        if val.shape[0] == 1 and hi - lo > 1:
            # f does not depend on the leading angle: one evaluation, weight = sum of this
            # rank's share of the rule
            part = _contract_dense(c, val, [_sum_weight(w_list[0][lo_rank:hi_rank])] + w_list[1:], None)
            out = part if out is None else out + part
            break
        part_w = [w_list[0][lo:hi]] + w_list[1:]
        out = _contract_dense(c, val, part_w, out)  # accumulates in place (new tensor under autograd)
        lo = hi
    if out is None:  # a rank with an empty share
        probe = f({k: (v[:2] if k == nodes[0] else v) for k, v in xs.items()})
        out = torch.zeros(probe.shape[c.s_ndim :], dtype=torch.promote_types(probe.dtype, w_list[0].dtype), device=device)
    return _finish(out, group, sharded=sharded)


class IntegralGraph:
    """``integrate(...)`` captured ONCE in a CUDA graph and replayed: the kernels ``f`` launches, the
    fused contraction and — for an integral sharded over ``group`` — the NCCL all-reduce of the
    partial sums all become one graph launch.

    What it is for: an integral that is evaluated again and again (a parameter sweep, an
    iteration) costs ``integrate`` a dozen launches and their host work per call, which at 8 GPUs
    (21 M nodes per rank of BASELINE's C5 grid) is more than the kernels themselves take.  ``f``
    must be capturable (device tensors in, no host synchronisation) and must read whatever
    changes between evaluations from tensors that are updated IN PLACE::

        k = torch.tensor([...], device="cuda", dtype=torch.float64)
        g = us.IntegralGraph(c, lambda s: torch.exp(c.to_cartesian_dot(s, k)), False, 96, xp=torch,
                             dtype=torch.float64, group=pg)
        for new_k in sweep:
            k.copy_(new_k)
            value = g()          # the graph's output buffer: valid until the next replay

    Same arguments and result as :func:`integrate` (dense or separable).  Every rank of ``group``
    must construct and replay the object alike — a replay is a collective call.
    """

    def __init__(
        self,
        c: SphericalCoordinates[Any, Any],
        f: Callable[[Mapping[Any, torch.Tensor]], Any] | Mapping[Any, torch.Tensor] | torch.Tensor,
        does_f_support_separation_of_variables: bool,
        n: int,
        *,
        xp: Any,
        device: Any | None = None,
        dtype: Any | None = None,
        slab_bytes: int | None = None,
        group: Any | None = None,
        warmup: int = 2,
    ) -> None:
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device())
        self.device = torch.device(device)
        run = lambda: integrate(  # noqa: E731
            c, f, does_f_support_separation_of_variables, n, xp=xp, device=self.device, dtype=dtype, slab_bytes=slab_bytes, group=group
        )
        with torch.cuda.device(self.device), torch.no_grad():
            # eager passes on a side stream first: rule tables, plan, scratch and NCCL's own buffers
            # must exist before the capture begins
            side = torch.cuda.Stream(self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):
                for _ in range(max(1, warmup)):
                    run()
            torch.cuda.current_stream(self.device).wait_stream(side)
            torch.cuda.synchronize(self.device)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.result = run()

    def __call__(self) -> Any:
        self.graph.replay()
        return self.result


def shard_rows(n_rows: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous share ``[lo, hi)`` of the leading grid axis owned by ``rank`` (may be empty)."""
    return (n_rows * rank) // world, (n_rows * (rank + 1)) // world


def _finish(result: torch.Tensor, group: Any | None, *, sharded: bool) -> torch.Tensor:
    if sharded:
        import torch.distributed as dist

        if result.requires_grad:
            raise RuntimeError("integrate(group=...): the all-reduce of a sharded integral is not differentiable; reduce the gradients yourself")

        dist.all_reduce(result, op=dist.ReduceOp.SUM, group=group)
    return result


def _integrate_separable(
    c: SphericalCoordinates[Any, Any], val: Mapping[Any, torch.Tensor], ws: Mapping[Any, torch.Tensor]
) -> dict[Any, torch.Tensor]:
    result: dict[Any, torch.Tensor] = {}
    todo_nodes, todo_vals, todo_w = [], [], []
    for node in c.s_nodes:
        value = val[node]
        w = ws[node]
        if value.shape[0] == 1:
            result[node] = value[0, ...] * w.sum()
        elif value.shape[0] == w.numel():
            result[node] = None  # type: ignore[assignment]  (keeps the reference's key order)
            todo_nodes.append(node)
            todo_vals.append(value)
            todo_w.append(w)
        else:
            raise ValueError(
                f"shape mismatch for {node}: values have {value.shape[0]} rows, the rule has {w.numel()} points"
            )
    if todo_nodes:
        for node, o in zip(todo_nodes, separable_reduce(todo_vals, todo_w)):
            result[node] = o
    return result
