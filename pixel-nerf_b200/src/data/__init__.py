"""Pass-through of the reference's `data` package (src/data/__init__.py: `get_split_dataset`, dataset adapters).

Dataset I/O is outside the render hot path, so nothing is re-implemented: this package's
search path IS the reference's `src/data` directory and its `__init__` is executed here, so `from data import
get_split_dataset` (eval/gen_video.py:14, train/train.py:15) gives the reference's own function."""
import _pnr_refpath

_ref_dir = _pnr_refpath.ref_src("data")
if _ref_dir is None:
    _pnr_refpath.need("the `data` package")
__path__.insert(0, _ref_dir)
with open(_ref_dir + "/__init__.py") as _f:
    exec(compile(_f.read(), _ref_dir + "/__init__.py", "exec"), globals())
