"""Drop-in shim: ``from tools import hierarchy`` / ``from tools.hierarchy import ...``
(model/multiscale_HSD.py:12, tests/robust_test/main.py:9 of the reference) resolve to the
B200-native modules.  The reference's remaining ``tools`` modules (evaluate, dataloader,
label, SIR, visualize, ... — downstream / upstream of the hot path, not rebuilt here) stay
importable from a reference checkout named by ``$HSD_REFERENCE_ROOT``: its ``tools/``
directory is appended to this package's search path, so ``from tools import evaluate``
(main.py:7) finds the reference's file while ``tools.hierarchy`` is the GPU one."""
import os as _os
import sys as _sys

from hsd_b200.tools import hierarchy, metrics, rw, util  # noqa: F401
from hsd_b200.tools.rw import save_vectors_dict  # noqa: F401
from hsd_b200.tools.hierarchy import *  # noqa: F401,F403
from hsd_b200.tools.util import *  # noqa: F401,F403

_sys.modules[__name__ + ".hierarchy"] = hierarchy
_sys.modules[__name__ + ".metrics"] = metrics
_sys.modules[__name__ + ".util"] = util
_sys.modules[__name__ + ".rw"] = rw

_ref = _os.environ.get("HSD_REFERENCE_ROOT")
if _ref and _os.path.isdir(_os.path.join(_ref, "tools")):
    __path__.append(_os.path.join(_ref, "tools"))
name = "tools"
