"""Stand-in for array_api_extra: only the two helpers the reference's hot path uses."""
import numpy as _np

broadcast_shapes = _np.broadcast_shapes


def isclose(a, b, *, rtol=1e-5, atol=1e-8, equal_nan=False, xp=None):
    return _np.isclose(a, b, rtol=rtol, atol=atol, equal_nan=equal_nan)
