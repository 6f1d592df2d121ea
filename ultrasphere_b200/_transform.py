"""Host glue of the transforms: dict-of-tensors ⇄ pointer tables ⇄ one kernel launch.

Semantics follow /root/reference/src/ultrasphere/_coordinates.py:555-579 (``from_cartesian``) and
:636-656 (``to_cartesian``): key orders, per-entry broadcast shapes, ``r`` defaulting to 1,
``KeyError`` for a missing node, array input indexed by leaf id.  The array work is done by
``us_to_cartesian`` / ``us_from_cartesian`` (include/ultrasphere_b200.h) on the caller's current
CUDA stream; outputs are fresh tensors that never alias the inputs.
"""
from __future__ import annotations

import ctypes as C
from collections.abc import Mapping, Sequence
from typing import Any

import torch

from . import _native
from ._coordinates import SphericalCoordinates
from ._plan import CompiledTree, plan_for

_DTYPE_CODE = {torch.float32: _native.US_F32, torch.float64: _native.US_F64}


def _needs_grad(values: Sequence[Any]) -> bool:
    return torch.is_grad_enabled() and any(isinstance(v, torch.Tensor) and v.requires_grad for v in values)


def _require_cuda_tensors(values: Sequence[Any], what: str, *, allow_grad: bool = False) -> torch.device:
    device: torch.device | None = None
    for v in values:
        if not isinstance(v, torch.Tensor):
            raise TypeError(
                f"{what}: expected CUDA torch.Tensor values, got {type(v).__name__}; "
                "ultrasphere_b200 has no NumPy / CPU path"
            )
        if v.device.type != "cuda":
            raise RuntimeError(
                f"{what}: tensor on {v.device}; ultrasphere_b200 runs on CUDA devices only (no CPU fallback)"
            )
        if not allow_grad and v.requires_grad and torch.is_grad_enabled():
            raise RuntimeError(
                f"{what}: an input requires grad, but the CUDA kernels have no backward pass; "
                "call under torch.no_grad() or detach() the inputs (silently dropping the graph would be worse)"
            )
        if device is None:
            device = v.device
        elif v.device != device:
            raise RuntimeError(f"{what}: tensors on different devices ({device} and {v.device})")
    assert device is not None
    return device


def broadcast_shapes(*shapes: Sequence[int]) -> tuple[int, ...]:
    """NumPy broadcasting of plain integer shapes.  ``torch.broadcast_shapes`` goes through the
    symbolic-shape machinery and costs ~35 us for five operands; this costs ~2."""
    nd = 0
    for sh in shapes:
        if len(sh) > nd:
            nd = len(sh)
    out = [1] * nd
    for sh in shapes:
        off = nd - len(sh)
        for i, e in enumerate(sh):
            if e != 1:
                j = off + i
                if out[j] == 1:
                    out[j] = e
                elif out[j] != e:
                    raise RuntimeError(f"shapes {[tuple(x) for x in shapes]} are not broadcastable")
    return tuple(out)


def _compute_dtype(tensors: Sequence[torch.Tensor]) -> torch.dtype:
    dt = tensors[0].dtype
    for t in tensors[1:]:
        if t.dtype != dt:
            dt = torch.promote_types(dt, t.dtype)
    if not dt.is_floating_point:
        if dt.is_complex:
            raise TypeError("complex coordinates are not supported")
        dt = torch.get_default_dtype()
    if dt not in _DTYPE_CODE:
        raise TypeError(f"unsupported dtype {dt}: the kernels compute in float32 or float64")
    return dt


class _Geometry:
    """Launch index space plus, for every operand, its dense-broadcast mask."""

    __slots__ = ("full_shape", "n", "shape", "masks", "tensors")

    def __init__(self, tensors: list[torch.Tensor | None], full_shape: tuple[int, ...]):
        self.full_shape = full_shape
        n = 1
        for e in full_shape:
            n *= e
        self.n = n
        live = [t for t in tensors if t is not None]
        # fast path: everything already dense over the full shape
        if all(t.shape == full_shape and t.is_contiguous() for t in live):
            self.shape = [n]
            self.masks = [0 if t is None else 1 for t in tensors]
            self.tensors = tensors
            return
        nd = len(full_shape)
        spans: list[list[bool] | None] = []
        fixed: list[torch.Tensor | None] = []
        for t in tensors:
            if t is None:
                spans.append(None)
                fixed.append(None)
                continue
            t, sp = _span_of(t, nd)
            spans.append(sp)
            fixed.append(t)
        # drop unit dims, merge neighbours that every operand treats alike
        dims = [d for d in range(nd) if full_shape[d] != 1]
        shape: list[int] = []
        groups: list[list[int]] = []
        for d in dims:
            if groups and all(sp is None or sp[groups[-1][-1]] == sp[d] for sp in spans):
                groups[-1].append(d)
                shape[-1] *= full_shape[d]
            else:
                groups.append([d])
                shape.append(full_shape[d])
        if not shape:
            shape, groups = [1], [[]]
        if len(shape) > _native.US_MAX_DIMS:
            # too many independent broadcast axes for the parameter block: densify on the device
            fixed = [None if t is None else t.expand(full_shape).contiguous() for t in fixed]
            self.shape = [n]
            self.masks = [0 if t is None else 1 for t in fixed]
            self.tensors = fixed
            return
        masks = []
        for sp in spans:
            m = 0
            if sp is not None:
                for gi, g in enumerate(groups):
                    if g and sp[g[0]]:
                        m |= 1 << gi
            masks.append(m)
        self.shape = shape
        self.masks = masks
        self.tensors = fixed


def _span_of(t: torch.Tensor, nd: int) -> tuple[torch.Tensor, list[bool]]:
    """Which launch dims does ``t`` really vary along, and is it dense over exactly those?"""
    pad = nd - t.dim()
    for _ in range(2):
        sp = [False] * nd
        expect = 1
        dense = True
        for d in range(t.dim() - 1, -1, -1):
            size, stride = t.shape[d], t.stride(d)
            if size == 1 or stride == 0:
                continue
            sp[pad + d] = True
            if stride != expect:
                dense = False
                break
            expect *= size
        if dense:
            return t, sp
        t = t.contiguous()  # arbitrary strided view: one device-side copy, then dense
    raise AssertionError("contiguous tensor must be dense")


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)
if _raw_stream is None:  # older torch: same value through the public API

    def _raw_stream(index: int) -> int:  # type: ignore[no-redef]
        return torch.cuda.current_stream(index).cuda_stream


# the calling thread's current CUDA device without the Python-level wrapper (0.1 us instead of 0.4)
_current_device = getattr(torch._C, "_cuda_getDevice", torch.cuda.current_device)


def _stream_ptr(device: torch.device) -> int:
    """``cudaStream_t`` of the caller's current stream on ``device`` (the raw getter skips building a
    ``torch.cuda.Stream`` object: 0.3 us instead of 3)."""
    if _raw_stream is not None:
        return int(_raw_stream(device.index if device.index is not None else torch.cuda.current_device()))
    return torch.cuda.current_stream(device).cuda_stream


class _NoGuard:
    def __enter__(self) -> None:
        return None

    def __exit__(self, *exc: Any) -> None:
        return None


_NO_GUARD = _NoGuard()


def device_guard(device: torch.device) -> Any:
    """Make ``device`` current for the launch (the library launches on the calling thread's
    current device); free when it already is."""
    if device.index is None or device.index == torch.cuda.current_device():
        return _NO_GUARD
    return torch.cuda.device(device)


_workspaces: dict[tuple[int, int], torch.Tensor] = {}


def _workspace_for(device: torch.device, nbytes: int) -> torch.Tensor:
    """Per-(device, stream) scratch owned by the shim (the library never allocates); grown on
    demand and reused, which is safe because every use is ordered on that one stream."""
    key = (device.index if device.index is not None else torch.cuda.current_device(), _stream_ptr(device))
    buf = _workspaces.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(nbytes, 1 << 16), dtype=torch.uint8, device=device)
        _workspaces[key] = buf
    return buf


def _ptr(t: torch.Tensor | None) -> int | None:
    return None if t is None else t.data_ptr()


# --------------------------------------------------------------------------------------
# autograd: dense forward launches + the backward kernels (us_*_backward)
# --------------------------------------------------------------------------------------
def _rows(t: torch.Tensor, count: int) -> Any:
    step = (t.numel() // count) * t.element_size() if count else 0
    return _native.void_p_array([t.data_ptr() + i * step for i in range(count)])


class _ToCartesianFn(torch.autograd.Function):
    """``(c_ndim, *shape)`` from dense equal-shape angles (s_nodes order) and an optional r."""

    @staticmethod
    def forward(ctx: Any, c: Any, has_r: bool, *ops: torch.Tensor) -> torch.Tensor:
        ct = plan_for(c)
        angles, r = (ops[:-1], ops[-1]) if has_r else (ops, None)
        first = angles[0]
        n = first.numel()
        out = torch.empty((ct.n_leaves, *first.shape), dtype=first.dtype, device=first.device)
        if n:
            with device_guard(first.device):
                status = _native.lib().us_to_cartesian(
                    C.byref(ct.plan), _DTYPE_CODE[first.dtype], n, 1, _native.i64_array([n]),
                    _native.void_p_array([a.data_ptr() for a in angles]), _native.u32_array([1] * ct.n_nodes),
                    _ptr(r), 1, _rows(out, ct.n_leaves), _stream_ptr(first.device),
                )
            _native.check(status, "us_to_cartesian")
        ctx.save_for_backward(*ops)
        ctx.c, ctx.has_r = c, has_r
        return out

    @staticmethod
    def backward(ctx: Any, grad_out: torch.Tensor) -> Any:
        ops = ctx.saved_tensors
        ct = plan_for(ctx.c)
        angles, r = (ops[:-1], ops[-1]) if ctx.has_r else (ops, None)
        first = angles[0]
        n = first.numel()
        g = grad_out.contiguous()
        grad_angles = torch.empty((ct.n_nodes, *first.shape), dtype=first.dtype, device=first.device)
        grad_r = torch.empty_like(first) if ctx.has_r else None
        if n:
            with device_guard(first.device):
                status = _native.lib().us_to_cartesian_backward(
                    C.byref(ct.plan), _DTYPE_CODE[first.dtype], n, _native.void_p_array([a.data_ptr() for a in angles]),
                    _ptr(r), _rows(g, ct.n_leaves), _rows(grad_angles, ct.n_nodes), _ptr(grad_r), _stream_ptr(first.device),
                )
            _native.check(status, "us_to_cartesian_backward")
        grads = list(grad_angles.unbind(0)) + ([grad_r] if ctx.has_r else [])
        return (None, None, *grads)


class _FromCartesianFn(torch.autograd.Function):
    """``(1 + s_ndim, *shape)`` — r, then the angles in s_nodes order — from dense equal-shape leaves."""

    @staticmethod
    def forward(ctx: Any, c: Any, *leaves: torch.Tensor) -> torch.Tensor:
        ct = plan_for(c)
        first = leaves[0]
        n = first.numel()
        out = torch.empty((1 + ct.n_nodes, *first.shape), dtype=first.dtype, device=first.device)
        if n:
            rows = _rows(out, 1 + ct.n_nodes)
            with device_guard(first.device):
                status = _native.lib().us_from_cartesian(
                    C.byref(ct.plan), _DTYPE_CODE[first.dtype], n, 1, _native.i64_array([n]),
                    _native.void_p_array([x.data_ptr() for x in leaves]), _native.u32_array([1] * ct.n_leaves),
                    rows[0], _native.void_p_array([rows[1 + s] for s in range(ct.n_nodes)]), _stream_ptr(first.device),
                )
            _native.check(status, "us_from_cartesian")
        ctx.save_for_backward(*leaves)
        ctx.c = c
        return out

    @staticmethod
    def backward(ctx: Any, grad_out: torch.Tensor) -> Any:
        leaves = ctx.saved_tensors
        ct = plan_for(ctx.c)
        first = leaves[0]
        n = first.numel()
        g = grad_out.contiguous()
        grad_x = torch.empty((ct.n_leaves, *first.shape), dtype=first.dtype, device=first.device)
        if n:
            scratch = torch.empty((2 * ct.n_nodes, n), dtype=first.dtype, device=first.device)
            rows = _rows(g, 1 + ct.n_nodes)
            with device_guard(first.device):
                status = _native.lib().us_from_cartesian_backward(
                    C.byref(ct.plan), _DTYPE_CODE[first.dtype], n, _native.void_p_array([x.data_ptr() for x in leaves]),
                    rows[0], _native.void_p_array([rows[1 + s] for s in range(ct.n_nodes)]), _rows(grad_x, ct.n_leaves),
                    scratch.data_ptr(), _stream_ptr(first.device),
                )
            _native.check(status, "us_from_cartesian_backward")
        return (None, *grad_x.unbind(0))


def _dense_operands(tensors: list[torch.Tensor], dtype: torch.dtype) -> list[torch.Tensor]:
    """Broadcast to one shape with differentiable torch ops: autograd sums the gradients back."""
    full = broadcast_shapes(*[t.shape for t in tensors])
    return [t.to(dtype).expand(full).contiguous() for t in tensors]


def _to_cartesian_with_grad(c: SphericalCoordinates[Any, Any], angles_in: list[Any], r_in: Any, as_array: bool) -> Any:
    r_is_tensor = isinstance(r_in, torch.Tensor)
    tensors = angles_in + ([r_in] if r_is_tensor else [])
    device = _require_cuda_tensors(tensors, "to_cartesian", allow_grad=True)
    dtype = _compute_dtype(tensors)
    if not r_is_tensor and r_in != 1:
        tensors = tensors + [torch.full((), float(r_in), dtype=dtype, device=device)]
    has_r = len(tensors) > len(angles_in)
    out = _ToCartesianFn.apply(c, has_r, *_dense_operands(tensors, dtype))
    return out if as_array else dict(zip(c.c_nodes, out.unbind(0)))


def _from_cartesian_with_grad(c: SphericalCoordinates[Any, Any], leaves_in: list[Any]) -> dict[Any, Any]:
    _require_cuda_tensors(leaves_in, "from_cartesian", allow_grad=True)
    dtype = _compute_dtype(leaves_in)
    out = _FromCartesianFn.apply(c, *_dense_operands(leaves_in, dtype))
    rows = out.unbind(0)
    result: dict[Any, Any] = {"r": rows[0]}
    for s in range(c.s_ndim - 1, -1, -1):
        result[c.s_nodes[s]] = rows[1 + s]
    return result


# --------------------------------------------------------------------------------------
# to_cartesian
# --------------------------------------------------------------------------------------
def to_cartesian(c: SphericalCoordinates[Any, Any], spherical: Mapping[Any, Any], *, as_array: bool = False) -> Any:
    angles_in = [spherical[k] for k in c.s_nodes]  # KeyError for a missing angle
    r_in = spherical.get("r", 1)
    if angles_in:
        # the common dense case first: it declines (None) anything that needs promotion, broadcasting or autograd
        fast = _to_cartesian_dense(c, angles_in, r_in, type(r_in) is torch.Tensor, as_array)
        if fast is not None:
            return fast
        if _needs_grad(angles_in + [r_in]):
            # differentiable path: every entry has the full broadcast shape (values as in the reference)
            return _to_cartesian_with_grad(c, angles_in, r_in, as_array)
    if c.s_ndim == 0:
        # a single leaf that is also the root: x = r  
        if as_array:
            if not isinstance(r_in, torch.Tensor):
                raise TypeError("to_cartesian(as_array=True) on a 0-angle tree needs a tensor r")
            return torch.stack([r_in], dim=0)
        return {c.c_nodes[0]: r_in}
    r_is_tensor = isinstance(r_in, torch.Tensor)
    device = _require_cuda_tensors(angles_in + ([r_in] if r_is_tensor else []), "to_cartesian")
    dtype = _compute_dtype(angles_in + ([r_in] if r_is_tensor else []))
    if not r_is_tensor:
        if isinstance(r_in, complex):
            raise TypeError("complex r is not supported")
        if r_in == 1:
            r_in = None
        else:
            r_in = torch.full((), float(r_in), dtype=dtype, device=device)
    angles = [a if a.dtype == dtype else a.to(dtype) for a in angles_in]
    r = None if r_in is None else (r_in if r_in.dtype == dtype else r_in.to(dtype))
    ct = plan_for(c)

    r_shape = tuple(r.shape) if r is not None else ()
    shapes = [tuple(a.shape) for a in angles]
    same = all(s == shapes[0] for s in shapes) and (r is None or r_shape == shapes[0] or r_shape == ())
    if same or as_array:
        full = shapes[0] if same else broadcast_shapes(r_shape, *shapes)
        out = _launch_to(ct, device, dtype, angles, r, full, list(range(ct.n_leaves)))
        if as_array:
            return out
        return {leaf: out[i] for i, leaf in enumerate(c.c_nodes)}
    # heterogeneous shapes: every leaf keeps the broadcast shape of r and its own ancestors
    leaf_shapes = [
        broadcast_shapes(r_shape, *[shapes[a] for a in ct.leaf_ancestors[l]]) for l in range(ct.n_leaves)
    ]
    result: dict[int, torch.Tensor] = {}
    for shape in dict.fromkeys(leaf_shapes):
        members = [l for l in range(ct.n_leaves) if leaf_shapes[l] == shape]
        out = _launch_to(ct, device, dtype, angles, r, tuple(shape), members)
        for j, l in enumerate(members):
            result[l] = out[j]
    return {leaf: result[i] for i, leaf in enumerate(c.c_nodes)}


def _to_cartesian_dense(
    c: SphericalCoordinates[Any, Any], angles: list[Any], r: Any, r_is_tensor: bool, as_array: bool
) -> Any:
    """The common case with the least host work: every operand a dense CUDA tensor of one shape
    and one float dtype (r absent, 1, or such a tensor too).  Returns None if the case is not this
    one; the general path then validates, promotes and broadcasts.  One ``torch.empty``, one call
    of ``us_to_cartesian_rows`` (outputs = rows of the allocation: no pointer or mask tables)."""
    first = angles[0]
    if type(first) is not torch.Tensor or not first.is_cuda:
        return None
    dtype, shape, device = first.dtype, first.shape, first.device
    code = _DTYPE_CODE.get(dtype)
    if code is None:
        return None
    if r_is_tensor:
        ops = angles + [r]
    elif r == 1:
        ops = angles
    else:
        return None
    grad = torch.is_grad_enabled()
    for t in ops:
        if (
            type(t) is not torch.Tensor
            or t.dtype != dtype
            or t.shape != shape
            or t.device != device
            or not t.is_contiguous()
            or (grad and t.requires_grad)
        ):
            return None
    ct = c._plan or plan_for(c)
    n = first.numel()
    out = first.new_empty((ct.n_leaves, *shape))
    if n:
        ptrs = ct.angle_ptrs  # reused table: filled and consumed under the GIL, before the call returns
        for i, a in enumerate(angles):
            ptrs[i] = a.data_ptr()
        if device.index != _current_device():
            with torch.cuda.device(device):
                status = _native.lib().us_to_cartesian_rows(
                    ct.plan_ref, code, n, ptrs, r.data_ptr() if r_is_tensor else None, out.data_ptr(), n * out.element_size(), _raw_stream(device.index)
                )
        else:
            status = ct.to_rows(
                ct.plan_ref, code, n, ptrs, r.data_ptr() if r_is_tensor else None, out.data_ptr(), n * out.element_size(), _raw_stream(device.index)
            )
        if status:
            _native.check(status, "us_to_cartesian")
    if as_array:
        return out
    return dict(zip(c.c_nodes, out.unbind(0)))


def _launch_to(
    ct: CompiledTree,
    device: torch.device,
    dtype: torch.dtype,
    angles: list[torch.Tensor],
    r: torch.Tensor | None,
    full: tuple[int, ...],
    members: list[int],
) -> torch.Tensor:
    """One launch over ``full``; returns a ``(len(members), *full)`` tensor (rows = member leaves)."""
    needed = set()
    for l in members:
        needed.update(ct.leaf_ancestors[l])
    # angles that feed none of the requested leaves are read at element 0 and their value unused
    ops: list[torch.Tensor | None] = [a if i in needed else None for i, a in enumerate(angles)]
    geo = _Geometry(ops + [r], full)
    out = torch.empty((len(members), *full), dtype=dtype, device=device)
    if geo.n == 0:
        return out
    in_ptrs = []
    for i, t in enumerate(geo.tensors[:-1]):
        in_ptrs.append(angles[i].data_ptr() if t is None else t.data_ptr())
    out_ptrs: list[int | None] = [None] * ct.n_leaves
    row_bytes = geo.n * out.element_size()
    base = out.data_ptr()
    for j, l in enumerate(members):
        out_ptrs[l] = base + j * row_bytes
    rt = geo.tensors[-1]
    lib = _native.lib()
    shape_arr = _native.i64_array(geo.shape)
    mask_arr = _native.u32_array(geo.masks[:-1])
    # broadcast operands (e.g. the 1-D quadrature axes an integrand receives): scratch for their
    # sin/cos tables, so the grid kernel only looks up, multiplies and stores
    ws_ptr, ws_bytes = None, 0
    if len(geo.shape) > 1 or any(m != 1 for m in geo.masks[:-1]):
        ws_bytes = int(
            lib.us_to_cartesian_workspace_bytes(C.byref(ct.plan), _DTYPE_CODE[dtype], geo.n, len(geo.shape), shape_arr, mask_arr)
        )
        if ws_bytes:
            ws = _workspace_for(device, ws_bytes)
            ws_ptr, ws_bytes = ws.data_ptr(), ws.numel()
    with device_guard(device):
        status = lib.us_to_cartesian_ws(
            C.byref(ct.plan),
            _DTYPE_CODE[dtype],
            geo.n,
            len(geo.shape),
            shape_arr,
            _native.void_p_array(in_ptrs),
            mask_arr,
            _ptr(rt),
            geo.masks[-1],
            _native.void_p_array(out_ptrs),
            ws_ptr,
            ws_bytes,
            _stream_ptr(device),
        )
    _native.check(status, "us_to_cartesian")
    return out


def to_cartesian_dot(c: SphericalCoordinates[Any, Any], spherical: Mapping[Any, Any], k: Any) -> torch.Tensor:
    """``einsum("v,v...->...", k, to_cartesian(spherical, as_array=True))`` without materialising the
    Cartesian coordinates (``us_to_cartesian_dot``).  ``k``: ``(c_ndim,)`` or ``(m, c_ndim)``
    coefficients in ``c_nodes`` order (tensor on any device, or numbers); returns the broadcast
    shape of the inputs, with a leading ``m`` axis for a matrix ``k``."""
    angles_in = [spherical[key] for key in c.s_nodes]  # KeyError for a missing angle
    r_in = spherical.get("r", 1)
    r_is_tensor = isinstance(r_in, torch.Tensor)
    if c.s_ndim > 0 and _needs_grad(angles_in + [r_in] + [k]):
        # under autograd: the contraction of the differentiable transform (gradients are not the hot path)
        x = to_cartesian(c, spherical, as_array=True)
        coef = torch.as_tensor(k, device=x.device).to(x.dtype)
        if coef.dim() not in (1, 2) or coef.shape[-1] != c.c_ndim:
            raise ValueError(f"k must have shape ({c.c_ndim},) or (m, {c.c_ndim}), got {tuple(coef.shape)}")
        return torch.einsum("v,v...->..." if coef.dim() == 1 else "mv,v...->m...", coef, x)
    tensors = angles_in + ([r_in] if r_is_tensor else [])
    if c.s_ndim == 0:
        tensors = [r_in] if r_is_tensor else []
        if not tensors:
            raise TypeError("to_cartesian_dot on a 0-angle tree needs a tensor r")
    device = _require_cuda_tensors(tensors, "to_cartesian_dot")
    dtype = _compute_dtype(tensors)
    coef = torch.as_tensor(k, device=device).to(dtype)
    if coef.dim() not in (1, 2) or coef.shape[-1] != c.c_ndim:
        raise ValueError(f"k must have shape ({c.c_ndim},) or (m, {c.c_ndim}), got {tuple(coef.shape)}")
    single = coef.dim() == 1
    coef = coef.reshape(-1, c.c_ndim).contiguous()
    m = coef.shape[0]
    if c.s_ndim == 0:
        out = coef[:, 0].reshape((m,) + (1,) * r_in.dim()) * r_in.to(dtype)
        return out[0] if single else out
    if not r_is_tensor:
        if isinstance(r_in, complex):
            raise TypeError("complex r is not supported")
        r_in = None if r_in == 1 else torch.full((), float(r_in), dtype=dtype, device=device)
    angles = [a if a.dtype == dtype else a.to(dtype) for a in angles_in]
    r = None if r_in is None else (r_in if r_in.dtype == dtype else r_in.to(dtype))
    ct = plan_for(c)
    full = tuple(broadcast_shapes(*[t.shape for t in angles + ([r] if r is not None else [])]))
    geo = _Geometry(list(angles) + [r], full)
    out = torch.empty((m, *full), dtype=dtype, device=device)
    if geo.n and m:
        lib = _native.lib()
        shape_arr = _native.i64_array(geo.shape)
        mask_arr = _native.u32_array(geo.masks[:-1])
        ws_ptr, ws_bytes = None, 0
        if len(geo.shape) > 1 or any(mk != 1 for mk in geo.masks[:-1]):
            ws_bytes = int(
                lib.us_to_cartesian_workspace_bytes(C.byref(ct.plan), _DTYPE_CODE[dtype], geo.n, len(geo.shape), shape_arr, mask_arr)
            )
            if ws_bytes:
                ws = _workspace_for(device, ws_bytes)
                ws_ptr, ws_bytes = ws.data_ptr(), ws.numel()
        with device_guard(device):
            status = lib.us_to_cartesian_dot(
                C.byref(ct.plan),
                _DTYPE_CODE[dtype],
                geo.n,
                len(geo.shape),
                shape_arr,
                _native.void_p_array([t.data_ptr() for t in geo.tensors[:-1]]),
                mask_arr,
                _ptr(geo.tensors[-1]),
                geo.masks[-1],
                coef.data_ptr(),
                m,
                out.data_ptr(),
                ws_ptr,
                ws_bytes,
                _stream_ptr(device),
            )
        _native.check(status, "us_to_cartesian_dot")
    return out[0] if single else out


# --------------------------------------------------------------------------------------
# from_cartesian
# --------------------------------------------------------------------------------------
def _from_cartesian_dense(c: SphericalCoordinates[Any, Any], cartesian: Any) -> dict[Any, Any] | None:
    """Least host work for the two common inputs: one dense ``(c_ndim, *shape)`` CUDA tensor (one
    call of ``us_from_cartesian_rows``: base pointers and row pitches, no tables), or a mapping of
    dense equal-shape CUDA tensors.  None if the input is anything else."""
    ct_nodes = c.c_nodes
    grad = torch.is_grad_enabled()
    ct = c._plan or plan_for(c)
    if type(cartesian) is torch.Tensor:
        x = cartesian
        if (
            not x.is_cuda
            or x.dim() < 1
            or x.shape[0] != ct.n_leaves
            or not x.is_contiguous()
            or not c.is_c_keys_sorted_range
            or (grad and x.requires_grad)
        ):
            return None
        dtype, device = x.dtype, x.device
        code = _DTYPE_CODE.get(dtype)
        if code is None:
            return None
        n = x.numel() // ct.n_leaves
        out = torch.empty_like(x)  # r and s_ndim angles: as many rows as leaves (the cheapest way to allocate them)
        if n:
            row = n * x.element_size()
            if device.index != _current_device():
                with torch.cuda.device(device):
                    status = ct.from_rows(ct.plan_ref, code, n, x.data_ptr(), row, out.data_ptr(), row, _raw_stream(device.index))
            else:
                status = ct.from_rows(ct.plan_ref, code, n, x.data_ptr(), row, out.data_ptr(), row, _raw_stream(device.index))
            if status:
                _native.check(status, "us_from_cartesian")
    else:
        try:
            leaves = [cartesian[k] for k in ct_nodes]
        except (KeyError, IndexError, TypeError):
            return None
        first = leaves[0]
        if type(first) is not torch.Tensor or not first.is_cuda:
            return None
        dtype, shape, device = first.dtype, first.shape, first.device
        for t in leaves:
            if (
                type(t) is not torch.Tensor
                or t.dtype != dtype
                or t.shape != shape
                or t.device != device
                or not t.is_contiguous()
                or (grad and t.requires_grad)
            ):
                return None
        code = _DTYPE_CODE.get(dtype)
        if code is None:
            return None
        n = first.numel()
        out = torch.empty((ct.n_nodes + 1, *shape), dtype=dtype, device=device)
        if n:
            row_bytes = n * out.element_size()
            base = out.data_ptr()
            with device_guard(device):
                status = _native.lib().us_from_cartesian(
                    ct.plan_ref,
                    code,
                    n,
                    1,
                    _native.i64_array([n]),
                    _native.void_p_array([t.data_ptr() for t in leaves]),
                    ct.ones_leaves,
                    base,
                    _native.void_p_array([base + (1 + s) * row_bytes for s in range(ct.n_nodes)]),
                    _stream_ptr(device),
                )
            _native.check(status, "us_from_cartesian")
    rows = out.unbind(0)
    result: dict[Any, Any] = {"r": rows[0]}
    for node, s in ct.from_order:  # dict order: "r", then children before parents
        result[node] = rows[s]
    return result


def from_cartesian(c: SphericalCoordinates[Any, Any], cartesian: Any) -> dict[Any, Any]:
    if c.s_nodes:
        fast = _from_cartesian_dense(c, cartesian)
        if fast is not None:
            return fast
    if isinstance(cartesian, torch.Tensor) and c.is_c_keys_sorted_range and cartesian.dim() >= 1 and cartesian.shape[0] == c.c_ndim:
        # one unbind instead of c_ndim selects: under autograd its backward is a single stack, where
        # every select would zero-fill and accumulate a full-size gradient (measured: 5x faster backward)
        leaves_in = list(cartesian.unbind(0))
    else:
        leaves_in = [cartesian[k] for k in c.c_nodes]  # mapping, or an object indexed by leaf id
    if c.s_ndim == 0:
        return {"r": leaves_in[0]}
    if _needs_grad(leaves_in):
        return _from_cartesian_with_grad(c, leaves_in)
    device = _require_cuda_tensors(leaves_in, "from_cartesian")
    dtype = _compute_dtype(leaves_in)
    leaves = [x if x.dtype == dtype else x.to(dtype) for x in leaves_in]
    ct = plan_for(c)
    shapes = [tuple(x.shape) for x in leaves]
    order = list(reversed(range(ct.n_nodes)))  # dict order: "r", then children before parents
    if all(s == shapes[0] for s in shapes):
        out = _launch_from(ct, device, dtype, leaves, shapes[0], True, list(range(ct.n_nodes)))
        result: dict[Any, Any] = {"r": out[0]}
        for s in order:
            result[c.s_nodes[s]] = out[1 + s]
        return result
    full = tuple(broadcast_shapes(*shapes))
    node_shapes = [tuple(broadcast_shapes(*[shapes[l] for l in ct.angle_leaves[s]])) for s in range(ct.n_nodes)]
    got: dict[int, torch.Tensor] = {}
    r_out: torch.Tensor | None = None
    for shape in dict.fromkeys([full] + node_shapes):
        members = [s for s in range(ct.n_nodes) if node_shapes[s] == shape]
        want_r = shape == full and r_out is None
        out = _launch_from(ct, device, dtype, leaves, shape, want_r, members)
        if want_r:
            r_out = out[0]
        for j, s in enumerate(members):
            got[s] = out[(1 if want_r else 0) + j]
    result = {"r": r_out}
    for s in order:
        result[c.s_nodes[s]] = got[s]
    return result


def _launch_from(
    ct: CompiledTree,
    device: torch.device,
    dtype: torch.dtype,
    leaves: list[torch.Tensor],
    full: tuple[int, ...],
    want_r: bool,
    members: list[int],
) -> torch.Tensor:
    """One launch over ``full``; returns ``(want_r + len(members), *full)``: r first, then the
    member angles in the given order."""
    if want_r:
        needed = set(range(ct.n_leaves))
    else:
        needed = set()
        for s in members:
            needed.update(ct.angle_leaves[s])
    ops: list[torch.Tensor | None] = [x if i in needed else None for i, x in enumerate(leaves)]
    geo = _Geometry(ops, full)
    rows = (1 if want_r else 0) + len(members)
    out = torch.empty((rows, *full), dtype=dtype, device=device)
    if geo.n == 0:
        return out
    in_ptrs = [leaves[i].data_ptr() if t is None else t.data_ptr() for i, t in enumerate(geo.tensors)]
    row_bytes = geo.n * out.element_size()
    base = out.data_ptr()
    out_ptrs: list[int | None] = [None] * ct.n_nodes
    for j, s in enumerate(members):
        out_ptrs[s] = base + ((1 if want_r else 0) + j) * row_bytes
    with device_guard(device):
        status = _native.lib().us_from_cartesian(
            C.byref(ct.plan),
            _DTYPE_CODE[dtype],
            geo.n,
            len(geo.shape),
            _native.i64_array(geo.shape),
            _native.void_p_array(in_ptrs),
            _native.u32_array(geo.masks),
            base if want_r else None,
            _native.void_p_array(out_ptrs),
            _stream_ptr(device),
        )
    _native.check(status, "us_from_cartesian")
    return out


# --------------------------------------------------------------------------------------
# jacobian
# --------------------------------------------------------------------------------------
def jacobian(c: SphericalCoordinates[Any, Any], spherical: Mapping[Any, Any]) -> Any:
    """``r * prod cos(theta)^(S_cos+1) * sin(theta)^(S_sin+1)`` (exponent 0 for leaf children) in one
    fused launch (``us_jacobian``).  Shapes broadcast like ``to_cartesian(as_array=True)``."""
    from ._coordinates import get_child

    angles_in = [spherical[k] for k in c.s_nodes]
    r_in = spherical["r"] if "r" in spherical else 1
    if c.s_ndim == 0:
        return r_in
    r_is_tensor = isinstance(r_in, torch.Tensor)
    tensors = angles_in + ([r_in] if r_is_tensor else [])
    cos_pow, sin_pow = [], []
    for node in c.s_nodes:
        for kind, out in (("cos", cos_pow), ("sin", sin_pow)):
            child = get_child(c.G, node, kind)
            out.append(0 if c.G.out_degree(child) == 0 else int(c.S[child]) + 1)
    if _needs_grad(tensors):
        # under autograd the product is spelled in torch ops (gradients are not the hot path)
        _require_cuda_tensors(tensors, "jacobian", allow_grad=True)
        val = r_in
        for a, cp, sp in zip(angles_in, cos_pow, sin_pow):
            val = val * torch.cos(a) ** cp * torch.sin(a) ** sp
        return val
    device = _require_cuda_tensors(tensors, "jacobian")
    dtype = _compute_dtype(tensors)
    if not r_is_tensor:
        r_in = None if r_in == 1 else torch.full((), float(r_in), dtype=dtype, device=device)
    angles = [a if a.dtype == dtype else a.to(dtype) for a in angles_in]
    r = None if r_in is None else (r_in if r_in.dtype == dtype else r_in.to(dtype))
    ct = plan_for(c)
    full = tuple(broadcast_shapes(*[t.shape for t in angles + ([r] if r is not None else [])]))
    geo = _Geometry(list(angles) + [r], full)
    out = torch.empty(full, dtype=dtype, device=device)
    if geo.n == 0:
        return out
    with device_guard(device):
        status = _native.lib().us_jacobian(
            C.byref(ct.plan),
            _DTYPE_CODE[dtype],
            geo.n,
            len(geo.shape),
            _native.i64_array(geo.shape),
            _native.void_p_array([t.data_ptr() for t in geo.tensors[:-1]]),
            _native.u32_array(geo.masks[:-1]),
            _ptr(geo.tensors[-1]),
            geo.masks[-1],
            (C.c_int32 * ct.n_nodes)(*cos_pow),
            (C.c_int32 * ct.n_nodes)(*sin_pow),
            out.data_ptr(),
            _stream_ptr(device),
        )
    _native.check(status, "us_jacobian")
    return out
