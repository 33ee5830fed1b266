from .hierarchy import *  # noqa: F401,F403
from .util import *  # noqa: F401,F403
from . import hierarchy, metrics, rw, util  # noqa: F401
from .rw import save_vectors_dict  # noqa: F401  (tools/multiscales.py:11 imports it from `tools`)

name = "tools"
