"""hsd_b200 — B200-native implementation of the HSD structural-distance hot path.

Importing the package loads ``libhsd_b200.so`` (hand-written sm_100a kernels
behind the C-ABI of ``include/hsd_b200.h``).  There is no CPU fallback: a
missing library raises at import, a missing GPU raises at first use.
"""
from . import _lib  # noqa: F401  (fails loudly when the library is not built)
from .graph import CSRGraph, DegreeOrder, powerlaw_graph  # noqa: F401

__all__ = ["CSRGraph", "DegreeOrder", "powerlaw_graph"]
__version__ = "0.1.0"
