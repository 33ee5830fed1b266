"""Drop-in shim: ``from model import HSD, MultiHSD`` (main.py:6 of the reference)
resolves to the B200-native classes in hsd_b200.model."""
from hsd_b200.model import HSD, MultiHSD, DynamicHSD, GraphWave  # noqa: F401
from hsd_b200.model import HSD as _m_HSD  # noqa: F401
import sys as _sys
import hsd_b200.model.HSD as _HSD_mod, hsd_b200.model.multiscale_HSD as _multi, \
    hsd_b200.model.dynamic_HSD as _dyn, hsd_b200.model.GraphWave as _gw

# `from model.multiscale_HSD import MultiHSD` (tests/robust_test/main.py:11) and friends
_sys.modules[__name__ + ".HSD"] = _HSD_mod
_sys.modules[__name__ + ".multiscale_HSD"] = _multi
_sys.modules[__name__ + ".dynamic_HSD"] = _dyn
_sys.modules[__name__ + ".GraphWave"] = _gw
name = "model"
