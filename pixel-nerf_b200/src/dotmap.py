"""`dotmap` stand-in for images that lack the real package (train/train.py:20 `from dotmap import DotMap`).
If a real `dotmap` distribution is installed anywhere else on sys.path it is loaded instead of this file."""
import importlib.machinery
import importlib.util
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
_spec = importlib.machinery.PathFinder.find_spec(
    "dotmap", [p for p in sys.path if os.path.abspath(p or ".") != _here])
if _spec is not None:
    _real = importlib.util.module_from_spec(_spec)
    sys.modules[__name__] = _real
    _spec.loader.exec_module(_real)
    DotMap = _real.DotMap
else:
    from render.dotmap_compat import DotMap  # noqa: F401
