"""Import alias: ``import pp_b200`` -> the package directory ``3d-object-detection_b200``
(whose name is not a valid Python identifier).  ``pp_b200.x`` and ``3d-object-detection_b200.x``
are the SAME module objects (a meta-path finder maps the alias names onto the real modules)."""
import importlib
import importlib.abc
import importlib.util
import os
import sys

_REAL = "3d-object-detection_b200"
_ALIAS = __name__
_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)


class _AliasFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if fullname.startswith(_ALIAS + "."):
            return importlib.util.spec_from_loader(fullname, self)
        return None

    def create_module(self, spec):
        mod = importlib.import_module(_REAL + spec.name[len(_ALIAS):])
        self._real_spec = getattr(mod, "__spec__", None)
        return mod

    def exec_module(self, module):
        if getattr(self, "_real_spec", None) is not None:
            module.__spec__ = self._real_spec      # keep the real identity of the shared module


if not any(isinstance(f, _AliasFinder) for f in sys.meta_path):
    sys.meta_path.insert(0, _AliasFinder())
sys.modules[__name__] = importlib.import_module(_REAL)
