"""Drop-in for the reference's `util` package (src/util/__init__.py:1-2): helper functions at top level plus the
`args` sub-module.

The ray / pose / indexing helpers next to the render hot path are this package's own (`util/util.py`).  Every other
name of the reference's grab-bag `src/util/util.py` -- `cmap`, `quat_to_rot`, `get_image_to_tensor_balanced`,
`get_mask_to_tensor`, padding helpers, ... (outside the hot path) -- is passed through,
on first use, to the reference's unmodified file (located by `_pnr_refpath`)."""
import _pnr_refpath

from .util import *  # noqa: F401,F403
from . import args  # noqa: F401
from . import hocon  # noqa: F401

_ref_util = None


def _reference_util():
    global _ref_util
    if _ref_util is None:
        import importlib.util as _ilu
        path = _pnr_refpath.ref_src("util", "util.py")
        if path is None:
            _pnr_refpath.need("this `util` helper")
        spec = _ilu.spec_from_file_location(__name__ + "._reference_util", path)
        mod = _ilu.module_from_spec(spec)
        spec.loader.exec_module(mod)
        _ref_util = mod
    return _ref_util


def __getattr__(name):
    if name.startswith("__"):
        raise AttributeError(name)
    try:
        return getattr(_reference_util(), name)
    except ImportError as e:
        raise AttributeError(f"module 'util' has no attribute {name!r} ({e})") from None
