"""Stand-in: the reference only uses jacobi_poly.lgamma on Python floats (surface_area/volume)."""
from math import lgamma  # noqa: F401
