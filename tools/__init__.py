"""Drop-in shim: ``from tools import hierarchy`` / ``from tools.hierarchy import ...``
(model/multiscale_HSD.py:12, tests/robust_test/main.py:9 of the reference)."""
import sys as _sys
from hsd_b200.tools import hierarchy, metrics, rw, util  # noqa: F401
from hsd_b200.tools.rw import save_vectors_dict  # noqa: F401
from hsd_b200.tools.hierarchy import *  # noqa: F401,F403
from hsd_b200.tools.util import *  # noqa: F401,F403

_sys.modules[__name__ + ".hierarchy"] = hierarchy
_sys.modules[__name__ + ".metrics"] = metrics
_sys.modules[__name__ + ".util"] = util
_sys.modules[__name__ + ".rw"] = rw
name = "tools"
