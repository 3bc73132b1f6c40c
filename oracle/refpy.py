"""Import the reference's OWN Python modules, unmodified, from /root/reference (TEST INFRASTRUCTURE ONLY).

``load()`` returns ``(box_utils, cfg)``: the reference's ``utils/box_utils.py`` module object and its
``config.cfg``, or ``None`` where the reference checkout is absent (the GPU box).  The third-party imports the
reference makes resolve to the stand-ins in ``oracle/sdk_shim`` (pyquaternion, lyft_dataset_sdk, easydict) and
``data.pillars`` to ``oracle/_ref`` (the reference's ``data/pillars.cpp`` built against the Boost stand-in).

Only ``tests/`` and ``tests/golden/make_golden_targets.py`` use this.
"""
import importlib
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
_loaded = None


def available():
    return os.path.exists(os.path.join(REF, "utils", "box_utils.py"))


def load():
    """(box_utils module, cfg) of the reference, or None."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        return None
    from . import ref
    pillars = ref.load()
    if pillars is None:
        ref.build(REF)
        pillars = ref.load()
    if pillars is None:
        return None
    shim = os.path.join(_HERE, "sdk_shim")
    for p in (REF, shim):
        if p not in sys.path:
            sys.path.insert(0, p)
    saved = {k: sys.modules.get(k) for k in ("config", "data", "data.pillars", "utils", "utils.box_utils")}
    for k in saved:
        sys.modules.pop(k, None)
    try:
        import data as ref_data                      # /root/reference/data/__init__.py
        sys.modules["data.pillars"] = pillars        # the compiled reference module (oracle/_ref)
        ref_data.pillars = pillars
        cfgmod = importlib.import_module("config")   # /root/reference/config.py
        bu = importlib.import_module("utils.box_utils")
        assert os.path.realpath(bu.__file__) == os.path.join(REF, "utils", "box_utils.py"), bu.__file__
    finally:
        # leave the reference's top-level names out of sys.modules: they would shadow this repo's
        # own ``config`` / ``utils`` lookups in the importing process
        for k, v in saved.items():
            if v is not None:
                sys.modules[k] = v
            else:
                sys.modules.pop(k, None)
        for p in (REF,):
            if p in sys.path:
                sys.path.remove(p)
    _loaded = (bu, cfgmod.cfg)
    return _loaded
