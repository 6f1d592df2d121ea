"""Minimal stand-in for array_api_compat: NumPy namespace unless a torch tensor is passed."""
import numpy as _np


def array_namespace(*xs):
    for x in xs:
        if type(x).__module__.split(".")[0] == "torch":
            from . import torch as _t

            return _t
    from . import numpy as _n

    return _n


def to_device(x, device):
    return x
