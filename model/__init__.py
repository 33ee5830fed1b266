"""Drop-in shim: ``from model import HSD, MultiHSD`` (main.py:6 of the reference)
resolves to the B200-native classes in hsd_b200.model, and the submodule paths the
reference's own callers use keep working:

* ``from model.multiscale_HSD import MultiHSD`` (tests/robust_test/main.py:11),
  ``from model.HSD import HSD``;
* ``from model import GraphWave; GraphWave.GraphWave(graph);
  GraphWave.recommend_scale_range(...)`` (tests/graphwave_test/main.py:12,33-34) —
  the reference's package does not re-export the GraphWave class, so ``model.GraphWave``
  is the MODULE there and stays the module here.
"""
import importlib as _importlib
import sys as _sys

from hsd_b200.model import HSD, MultiHSD, DynamicHSD  # noqa: F401  (model/__init__.py:2-4)

# `hsd_b200.model` rebinds the attribute names HSD / GraphWave to the classes, so the
# modules have to come from importlib (sys.modules), not from attribute access
_sub = {n: _importlib.import_module("hsd_b200.model." + n)
        for n in ("HSD", "multiscale_HSD", "dynamic_HSD", "GraphWave")}
for _n, _m in _sub.items():
    _sys.modules[__name__ + "." + _n] = _m
GraphWave = _sub["GraphWave"]
multiscale_HSD = _sub["multiscale_HSD"]
dynamic_HSD = _sub["dynamic_HSD"]
name = "model"
