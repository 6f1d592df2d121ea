"""Load the UNMODIFIED reference hot path from /root/reference (build container only).

TEST INFRASTRUCTURE — never imported by the product package.

The reference is pure Python; its package ``__init__`` pulls in matplotlib (absent here), so
the four hot-path modules are imported directly under an empty ``ultrasphere`` package shell
(``_coordinates.py``, ``_creation.py``, ``_integral.py``, ``_random.py``), with the missing
third-party packages replaced by the stand-ins in ``oracle/stubs``.

`/root/reference` does not exist on the GPU box, so this module is used only
  * by ``oracle/gen_golden.py`` to write ``tests/golden/*`` (committed), and
  * by ``tests/test_oracle_vs_reference.py`` which skips when the mount is absent.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

REFERENCE_SRC = "/root/reference/src/ultrasphere"
_STUBS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "stubs")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_SRC, "_coordinates.py"))


def load() -> types.SimpleNamespace:
    """Return a namespace with the reference's modules: .coordinates .creation .integral .random"""
    if not available():
        raise RuntimeError("reference sources are not mounted at /root/reference")
    if _STUBS not in sys.path:
        sys.path.insert(0, _STUBS)
    if "ultrasphere" not in sys.modules:
        shell = types.ModuleType("ultrasphere")
        shell.__path__ = [REFERENCE_SRC]  # type: ignore[attr-defined]
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
