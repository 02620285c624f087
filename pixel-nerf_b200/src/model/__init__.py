"""Drop-in for the reference's `model` package (src/model/__init__.py).

Hot-path modules (`models`, `resnetfc`, `encoder`, `code`, `model_util`) are this package's own.  Sub-modules outside
the hot path -- `model.loss` (train/train.py:13 `from model import make_model, loss`), `model.mlp`,
`model.custom_encoder` -- are passed through to the reference's unmodified files by appending its `src/model`
directory to this package's search path (this package's files always win)."""
import _pnr_refpath

from .models import PixelNeRFNet

_ref_dir = _pnr_refpath.ref_src("model")
if _ref_dir is not None and _ref_dir not in __path__:
    __path__.append(_ref_dir)


def make_model(conf, *args, **kwargs):
    """Factory with the reference's signature (src/model/__init__.py:4-11)."""
    model_type = conf.get_string("type", "pixelnerf")
    if model_type != "pixelnerf":
        raise NotImplementedError("Unsupported model type", model_type)
    return PixelNeRFNet(conf, *args, **kwargs)


def __getattr__(name):
    if name in ("loss", "mlp", "custom_encoder"):
        if _ref_dir is None:
            _pnr_refpath.need(f"model.{name}")
        import importlib
        return importlib.import_module(f"{__name__}.{name}")
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
