"""ultrasphere_b200 — the B200-native hot path of ultrasphere.

Drop-in for the reference's transform and quadrature surface
(/root/reference/src/ultrasphere/__init__.py:2-15): the ``create_*`` factories,
``SphericalCoordinates`` with ``to_cartesian`` / ``from_cartesian``, ``roots`` and ``integrate``
(``xp=torch``).  All array work runs in hand-written sm_100a CUDA kernels behind the C ABI in
``include/ultrasphere_b200.h``; there is no CPU fallback.
"""
from ._coordinates import BranchingType, SphericalCoordinates, get_child, get_parent
from ._creation import (
    create_from_branching_types,
    create_hopf,
    create_polar,
    create_random,
    create_spherical,
    create_standard,
    create_standard_prime,
)
from ._integral import integrate, roots, separable_reduce, weighted_reduce
from ._random import random_ball
from . import _ops as ops  # registers torch.ops.ultrasphere_b200.*
from ._ops import plan_handle

__version__ = "0.1.0"

__all__ = [
    "BranchingType",
    "SphericalCoordinates",
    "create_from_branching_types",
    "create_hopf",
    "create_polar",
    "create_random",
    "create_spherical",
    "create_standard",
    "create_standard_prime",
    "get_child",
    "get_parent",
    "integrate",
    "ops",
    "plan_handle",
    "random_ball",
    "roots",
    "separable_reduce",
    "weighted_reduce",
]
