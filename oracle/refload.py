"""Load the UNMODIFIED reference hot path from /root/reference (build container only).

TEST INFRASTRUCTURE — never imported by the product package.

The reference is pure Python; its package ``__init__`` pulls in matplotlib (absent here), so
the four hot-path modules are imported directly under an empty ``ultrasphere`` package shell
(``_coordinates.py``, ``_creation.py``, ``_integral.py``, ``_random.py``), with the missing
third-party packages replaced by the stand-ins in ``oracle/stubs``.

`/root/reference` does not exist on the GPU box.  There the same four files are loaded from the
git-ignored copy ``oracle/_ref/ultrasphere`` that ``oracle/make_ref.py`` makes in the build
container (it ships with the ``gpurun`` snapshot).  Users of this module:
  * ``oracle/gen_golden.py`` writes ``tests/golden/*`` (committed) from the mounted reference,
  * ``tests/test_oracle_vs_reference.py`` re-checks the oracle bit-for-bit (skips without either copy),
  * ``bench.py --impl reference`` and its ``cpu_baseline`` leg time the reference itself.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
_MOUNT = "/root/reference/src/ultrasphere"
_COPY = os.path.join(_HERE, "_ref", "ultrasphere")
_STUBS = os.path.join(_HERE, "stubs")


def source_dir() -> str | None:
    """Directory holding the reference's modules: the mount if present, else the shipped copy."""
    # US_REF_IGNORE_MOUNT=1: exercise the shipped copy in the build container (tests)
    dirs = (_COPY,) if os.environ.get("US_REF_IGNORE_MOUNT") == "1" else (_MOUNT, _COPY)
    for d in dirs:
        if os.path.isfile(os.path.join(d, "_coordinates.py")):
            return d
    return None


REFERENCE_SRC = source_dir() or _MOUNT


def available() -> bool:
    return source_dir() is not None


def load() -> types.SimpleNamespace:
    """Return a namespace with the reference's modules: .coordinates .creation .integral .random"""
    src = source_dir()
    if src is None:
        raise RuntimeError("reference sources are neither mounted at /root/reference nor copied to oracle/_ref (oracle/make_ref.py)")
    if _STUBS not in sys.path:
        sys.path.insert(0, _STUBS)
    if "ultrasphere" not in sys.modules:
        shell = types.ModuleType("ultrasphere")
        shell.__path__ = [src]  # type: ignore[attr-defined]
        sys.modules["ultrasphere"] = shell
    mods = {}
    for name in ("_coordinates", "_creation", "_integral", "_random"):
        mods[name.lstrip("_")] = importlib.import_module(f"ultrasphere.{name}")
    shell = sys.modules["ultrasphere"]
    # what `import ultrasphere as us` exposes for the hot path (reference __init__.py:2-15)
    for attr in ("SphericalCoordinates", "BranchingType", "get_child", "get_parent"):
        setattr(shell, attr, getattr(mods["coordinates"], attr))
    for attr in (
        "create_from_branching_types",
        "create_hopf",
        "create_polar",
        "create_random",
        "create_spherical",
        "create_standard",
        "create_standard_prime",
    ):
        setattr(shell, attr, getattr(mods["creation"], attr))
    shell.integrate = mods["integral"].integrate
    shell.roots = mods["integral"].roots
    shell.random_ball = mods["random"].random_ball
    return types.SimpleNamespace(us=shell, **mods)
