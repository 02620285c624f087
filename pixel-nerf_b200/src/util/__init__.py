"""Drop-in for the reference's `util` package (src/util/__init__.py:1-2): helper functions
at top level plus the `args` sub-module."""
from .util import *  # noqa: F401,F403
from . import args  # noqa: F401
from . import hocon  # noqa: F401
