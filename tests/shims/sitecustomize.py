"""TEST SHIM (not product code): environment compatibility for running the UNMODIFIED reference (written for
torch 1.6 / torchvision 0.7) in this image.  Loaded automatically because tests put tests/shims on PYTHONPATH.

`src/data/data_util.py:4` imports `torchvision.transforms.functional_tensor`, which modern torchvision renamed to
`_functional_tensor`; alias it on demand.  Chains to the interpreter's own sitecustomize."""
import importlib
import importlib.abc
import importlib.util
import sys

_OLD, _NEW = "torchvision.transforms.functional_tensor", "torchvision.transforms._functional_tensor"


class _Alias(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, name, path, target=None):
        if name != _OLD:
            return None
        return importlib.util.spec_from_loader(name, self)

    def create_module(self, spec):
        return importlib.import_module(_NEW)

    def exec_module(self, module):
        pass


sys.meta_path.insert(0, _Alias())

try:   # the system-wide sitecustomize this file shadows
    _spec = importlib.util.spec_from_file_location("_system_sitecustomize", "/usr/lib/python3.12/sitecustomize.py")
    if _spec is not None:
        _m = importlib.util.module_from_spec(_spec)
        _spec.loader.exec_module(_m)
except Exception:
    pass
