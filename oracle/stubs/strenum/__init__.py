"""Stand-in for the `strenum` backport; Python 3.12 has enum.StrEnum."""
from enum import StrEnum  # noqa: F401
