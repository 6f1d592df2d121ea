"""ultrasphere_b200 — the B200-native hot path of ultrasphere.

Drop-in for the reference's transform and quadrature surface: the ``create_*`` factories,
``SphericalCoordinates`` with ``to_cartesian`` / ``from_cartesian``, ``roots``, ``integrate`` and
``random_ball`` (``xp=torch``).  All array work runs in hand-written sm_100a CUDA kernels behind the
C ABI in ``include/ultrasphere_b200.h``; there is no CPU fallback.
"""
from . import _coordinates as _co, _creation as _cr, _integral as _in, _ops as ops, _random as _ra

__version__ = "0.1.0"

# coordinate object and tree helpers
BranchingType, SphericalCoordinates = _co.BranchingType, _co.SphericalCoordinates
get_child, get_parent = _co.get_child, _co.get_parent
# factories
create_polar, create_spherical = _cr.create_polar, _cr.create_spherical
create_standard, create_standard_prime = _cr.create_standard, _cr.create_standard_prime
create_hopf, create_random = _cr.create_hopf, _cr.create_random
create_from_branching_types = _cr.create_from_branching_types
# quadrature, sampling, fused reductions
roots, integrate = _in.roots, _in.integrate
IntegralGraph = _in.IntegralGraph  # an extension: integrate() captured in a CUDA graph, NCCL all-reduce included
rule_backend, set_rule_backend = _in.rule_backend, _in.set_rule_backend
weighted_reduce, separable_reduce = _in.weighted_reduce, _in.separable_reduce
random_ball = _ra.random_ball
# torch.ops.ultrasphere_b200.* (registered by importing `ops`)
plan_handle = ops.plan_handle

__all__ = tuple(
    sorted(
        name
        for name, value in list(globals().items())
        if not name.startswith("_") and (callable(value) or name == "ops")
    )
)
