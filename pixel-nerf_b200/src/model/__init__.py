"""Drop-in for the reference's `model` package (src/model/__init__.py)."""
from .models import PixelNeRFNet


def make_model(conf, *args, **kwargs):
    """Factory with the reference's signature (src/model/__init__.py:4-11)."""
    model_type = conf.get_string("type", "pixelnerf")
    if model_type != "pixelnerf":
        raise NotImplementedError("Unsupported model type", model_type)
    return PixelNeRFNet(conf, *args, **kwargs)
