"""Type-only stand-in for `types-array-api` (the reference imports it for annotations)."""
from typing import Any

Array = Any
ArrayNamespaceFull = Any
