from .hierarchy import *  # noqa: F401,F403
from .util import *  # noqa: F401,F403
from . import hierarchy, metrics, util  # noqa: F401

name = "tools"
