"""`dotmap.DotMap` when installed, otherwise a minimal equivalent (attribute access,
auto-vivified children, toDict) -- the reference returns DotMaps from NeRFRenderer.forward
(src/render/nerf.py:278,313) and callers use both `out.fine.rgb` and `out.toDict()`."""
try:  # pragma: no cover
    from dotmap import DotMap
except ImportError:
    class DotMap(dict):
        def __init__(self, *args, **kwargs):
            super().__init__()
            for k, v in dict(*args, **kwargs).items():
                self[k] = DotMap(v) if isinstance(v, dict) and not isinstance(v, DotMap) else v

        def __getattr__(self, key):
            if key.startswith("__"):
                raise AttributeError(key)
            if key not in self:
                self[key] = DotMap()
            return self[key]

        def __setattr__(self, key, value):
            self[key] = value

        def __delattr__(self, key):
            del self[key]

        def toDict(self):
            return {k: (v.toDict() if isinstance(v, DotMap) else v) for k, v in self.items()}
