"""Loader for the reference's own PyTorch modules (TEST / BENCH-BASELINE INFRASTRUCTURE ONLY).

``load()`` returns ``(module, cfg, kind)``: the reference's ``model.model`` module (PPFeatureNet, PPScatter, PPModel
... of /root/reference/model/model.py) and its ``config.cfg``
  * in the build container from the sources where they lie under /root/reference (kind "source"),
  * on the GPU box from oracle/_ref/refpy/*.pycode -- the same files byte-compiled by ``make -C oracle refpy``
    (build outputs only: git-ignored, they travel with the snapshot like the compiled pillars module; kind
    "bytecode"),
or None when neither exists.  ``easydict`` resolves to oracle/sdk_shim.  Nothing under 3d-object-detection_b200/
imports this; bench.py uses it for the ``gpu_comparator`` leg and tests/test_gpu_integration.py for the call-site test.
"""
import importlib
import marshal
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
_loaded = None


def _exec_pycode(name, path, package=None):
    with open(path, "rb") as f:
        data = f.read()
    code = marshal.loads(data[16:])                  # PEP 552 header: magic, flags, mtime / hash, size
    mod = types.ModuleType(name)
    mod.__file__ = path
    if package is not None:
        mod.__package__ = package
    sys.modules[name] = mod
    exec(code, mod.__dict__)
    return mod


def load():
    global _loaded
    if _loaded is not None:
        return _loaded
    shim = os.path.join(_HERE, "sdk_shim")
    saved = {k: sys.modules.get(k) for k in ("config", "model", "model.model")}
    for k in saved:
        sys.modules.pop(k, None)
    sys.path.insert(0, shim)
    try:
        if os.path.exists(os.path.join(REF, "model", "model.py")):
            sys.path.insert(0, REF)
            try:
                cfgmod = importlib.import_module("config")
                mod = importlib.import_module("model.model")
            finally:
                sys.path.remove(REF)
            kind = "source"
        else:
            d = os.path.join(_HERE, "_ref", "refpy")
            if not os.path.exists(os.path.join(d, "model.model.pycode")):
                return None
            cfgmod = _exec_pycode("config", os.path.join(d, "config.pycode"))
            pkg = types.ModuleType("model")
            pkg.__path__ = []
            sys.modules["model"] = pkg
            mod = _exec_pycode("model.model", os.path.join(d, "model.model.pycode"), package="model")
            kind = "bytecode"
    finally:
        sys.path.remove(shim)
        for k, v in saved.items():               # keep the reference's top-level names out of sys.modules
            if v is not None:
                sys.modules[k] = v
            else:
                sys.modules.pop(k, None)
    _loaded = (mod, cfgmod.cfg, kind)
    return _loaded
