"""ctypes binding of ``libultrasphere_b200.so`` (C ABI declared in ``include/ultrasphere_b200.h``).

There is deliberately no fallback: if the shared library is missing or fails a call, the
product raises.  Build it with ``python -c "import __graft_entry__ as g; g.build()"`` or
``make -C ultrasphere_b200/csrc``.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Any

US_MAX_NODES = 256
US_MAX_LEAVES = US_MAX_NODES + 1
US_MAX_DIMS = 12
US_MAX_STACK = 64
US_MAX_AXES = 64
US_MAX_RULE_POINTS = 2048
US_F32, US_F64 = 0, 1
US_ABI_VERSION = 1

# ULTRASPHERE_B200_LIB=<path>: load another build of the library (A/B measurements of kernel variants)
LIB_PATH = os.environ.get("ULTRASPHERE_B200_LIB") or os.path.join(
    os.path.dirname(os.path.abspath(__file__)), "lib", "libultrasphere_b200.so"
)


class UsPlan(C.Structure):
    """Mirror of ``us_plan_t``."""

    _fields_ = [
        ("n_nodes", C.c_int32),
        ("n_leaves", C.c_int32),
        ("max_stack", C.c_int32),
        ("reserved", C.c_int32),
        ("node", C.c_uint64 * US_MAX_NODES),
    ]


class NativeError(RuntimeError):
    """A call into libultrasphere_b200 returned a non-zero status."""


_lib: C.CDLL | None = None

_PP = C.POINTER(C.c_void_p)
_PI64 = C.POINTER(C.c_int64)
_PU32 = C.POINTER(C.c_uint32)

_SIGNATURES: dict[str, tuple[Any, list[Any]]] = {
    "us_abi_version": (C.c_int, []),
    "us_version": (C.c_char_p, []),
    "us_last_error_string": (C.c_char_p, []),
    "us_device_count": (C.c_int, []),
    "us_set_tuning": (C.c_int, [C.c_char_p, C.c_int]),
    "us_plan_compile": (
        C.c_int,
        [C.c_int32, C.POINTER(C.c_uint8), C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(UsPlan)],
    ),
    "us_to_cartesian": (
        C.c_int,
        [C.POINTER(UsPlan), C.c_int, C.c_int64, C.c_int, _PI64, _PP, _PU32, C.c_void_p, C.c_uint32, _PP, C.c_void_p],
    ),
    "us_to_cartesian_workspace_bytes": (C.c_int64, [C.POINTER(UsPlan), C.c_int, C.c_int64, C.c_int, _PI64, _PU32]),
    "us_to_cartesian_ws": (
        C.c_int,
        [C.POINTER(UsPlan), C.c_int, C.c_int64, C.c_int, _PI64, _PP, _PU32, C.c_void_p, C.c_uint32, _PP, C.c_void_p, C.c_int64, C.c_void_p],
    ),
    "us_jacobian": (
        C.c_int,
        [C.POINTER(UsPlan), C.c_int, C.c_int64, C.c_int, _PI64, _PP, _PU32, C.c_void_p, C.c_uint32,
         C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_void_p, C.c_void_p],
    ),
    "us_from_cartesian": (
        C.c_int,
        [C.POINTER(UsPlan), C.c_int, C.c_int64, C.c_int, _PI64, _PP, _PU32, C.c_void_p, _PP, C.c_void_p],
    ),
    "us_from_cartesian_rows": (C.c_int, [C.POINTER(UsPlan), C.c_int, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p]),
    "us_to_cartesian_rows": (C.c_int, [C.POINTER(UsPlan), C.c_int, C.c_int64, _PP, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "us_weighted_reduce_scratch_bytes": (C.c_int64, [C.c_int64]),
    "us_weighted_reduce": (
        C.c_int,
        [C.c_int, C.c_int, _PI64, _PP, C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_void_p],
    ),
    "us_separable_reduce": (C.c_int, [C.c_int, C.c_int, _PI64, _PI64, _PP, _PP, _PP, C.c_void_p]),
    "us_random_ball": (C.c_int, [C.c_int, C.c_int, C.c_int64, C.c_uint64, C.c_int, _PP, C.c_void_p]),
    "us_random_uniform": (
        C.c_int,
        [C.c_int, C.c_int, C.c_int64, C.c_uint64, C.POINTER(C.c_double), C.POINTER(C.c_double), _PP, C.c_void_p],
    ),
    "us_to_cartesian_dot": (
        C.c_int,
        [C.POINTER(UsPlan), C.c_int, C.c_int64, C.c_int, _PI64, _PP, C.POINTER(C.c_uint32), C.c_void_p, C.c_uint32,
         C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p],
    ),
    "us_to_cartesian_backward": (C.c_int, [C.POINTER(UsPlan), C.c_int, C.c_int64, _PP, C.c_void_p, _PP, _PP, C.c_void_p, C.c_void_p]),
    "us_from_cartesian_backward": (
        C.c_int,
        [C.POINTER(UsPlan), C.c_int, C.c_int64, _PP, C.c_void_p, _PP, _PP, C.c_void_p, C.c_void_p],
    ),
    "us_quadrature_rules": (
        C.c_int,
        [C.c_int, C.POINTER(C.c_uint8), _PI64, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int32), _PP, _PP, C.c_void_p],
    ),
    "us_probe_fma_peak": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_double)]),
    "us_math_probe_f32": (C.c_int, [C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "us_math_probe_f64": (C.c_int, [C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
}

EXPORTED = tuple(_SIGNATURES)


def lib() -> C.CDLL:
    """Load (once) and return the shared library; raise loudly when it is not there."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: ultrasphere_b200 has no CPU or PyTorch fallback. "
                "Build it with `make -C ultrasphere_b200/csrc` (needs nvcc, targets sm_100a)."
            )
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        got = handle.us_abi_version()
        if got != US_ABI_VERSION:
            raise ImportError(f"{LIB_PATH}: ABI version {got}, this package expects {US_ABI_VERSION}")
        _lib = handle
    return _lib


def check(status: int, what: str) -> None:
    if status != 0:
        msg = lib().us_last_error_string().decode(errors="replace")
        if status in (-2, -3):
            raise ValueError(f"{what}: {msg}")
        raise NativeError(f"{what} failed with status {status}: {msg}")


def void_p_array(values: list[int | None]) -> Any:
    arr = (C.c_void_p * max(len(values), 1))()
    for i, v in enumerate(values):
        arr[i] = v
    return arr


def i64_array(values: list[int]) -> Any:
    return (C.c_int64 * max(len(values), 1))(*values)


def u32_array(values: list[int]) -> Any:
    return (C.c_uint32 * max(len(values), 1))(*values)
