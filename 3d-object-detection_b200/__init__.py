"""pp_b200 -- B200-native (sm_100a) PointPillars input path.

A from-scratch implementation of the per-sweep hot path of mr3543/3d-Object-Detection behind the
reference's own interfaces:

    pillars.create_pillars / pillars.make_ious   (data/pillars.cpp pybind11 module)
    model.PPFeatureNet / model.PPScatter         (model/model.py; forward and parameter gradients)
    loss.PPLoss                                  (model/loss.py; losses + gradients in one fused pass)
    box_utils.create_target                      (utils/box_utils.py)
    pipeline.InputPath                           (device-native batch entry points)

All compute is hand-written CUDA in libpp_b200.so (csrc/, C ABI in include/pp_b200.h); PyTorch
only provides device memory, streams and torch.distributed.  There is no CPU fallback.

The directory is named ``3d-object-detection_b200`` (not an identifier); import it as
``import pp_b200`` via the alias module at the repo root.
"""
from . import _lib, build, config, synth  # noqa: F401
from .config import PPConfig, cfg  # noqa: F401

__all__ = ["pillars", "model", "loss", "box_utils", "pipeline", "synth", "config", "PPConfig", "cfg"]


def __getattr__(name):
    # torch-dependent sub-modules are imported lazily so that build / symbol checks stay light
    if name in ("pillars", "model", "loss", "box_utils", "pipeline", "_runtime"):
        import importlib
        return importlib.import_module("." + name, __name__)
    raise AttributeError(name)
