"""Drop-in for the reference's `render` package (src/render/__init__.py)."""
from .nerf import NeRFRenderer  # noqa: F401
from .frames import render_frames  # noqa: F401
