"""ctypes wrapper over oracle/pp_oracle.c (TEST INFRASTRUCTURE ONLY, see oracle/__init__.py)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libpp_oracle.so")
_lib = None


def build(force=False):
    """Compile pp_oracle.c with gcc (flags in oracle/Makefile)."""
    src = os.path.join(_HERE, "pp_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "oracle"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        dp = ctypes.POINTER(ctypes.c_double)
        L.pp_oracle_create_pillars.restype = ctypes.c_int
        L.pp_oracle_create_pillars.argtypes = [
            ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p,
            ctypes.c_void_p, ctypes.c_int, ctypes.c_int] + [ctypes.c_double] * 9 + [
            ctypes.POINTER(ctypes.c_int64)]
        L.pp_oracle_iou.restype = ctypes.c_double
        L.pp_oracle_iou.argtypes = [dp, dp]
        L.pp_oracle_make_ious.restype = ctypes.c_int
        L.pp_oracle_make_ious.argtypes = [ctypes.c_void_p] * 5 + [ctypes.c_int64, ctypes.c_int64]
        _lib = L
    return _lib


def create_pillars(points, tensor, indices, max_points_per_pillar, max_pillars, x_step, y_step,
                   x_min, y_min, z_min, x_max, y_max, z_max, canvas_height):
    """Same positional signature and in-place semantics as the reference's
    ``pillars.create_pillars`` (data/pillars.cpp:236-249).  ``points`` may be any strided
    float64 [Npts, >=4] view.  Returns the number of pillars written."""
    assert points.dtype == np.float64 and points.ndim == 2 and points.shape[1] >= 4
    assert tensor.dtype == np.float64 and tensor.flags.c_contiguous
    assert indices.dtype == np.float64 and indices.flags.c_contiguous
    n = ctypes.c_int64(0)
    es = points.itemsize
    rc = lib().pp_oracle_create_pillars(
        points.ctypes.data, points.shape[0], points.strides[0] // es, points.strides[1] // es,
        tensor.ctypes.data, indices.ctypes.data, int(max_points_per_pillar), int(max_pillars),
        float(x_step), float(y_step), float(x_min), float(y_min), float(z_min), float(x_max),
        float(y_max), float(z_max), float(canvas_height), ctypes.byref(n))
    if rc != 0:
        raise RuntimeError("pp_oracle_create_pillars failed: %d" % rc)
    return int(n.value)


def iou(a_ring, g_ring):
    """IoU of one CCW anchor ring [4,2] and one CW GT ring [4,2] (data/pillars.cpp:132-172)."""
    a = np.ascontiguousarray(a_ring, dtype=np.float64)
    g = np.ascontiguousarray(g_ring, dtype=np.float64)
    dp = ctypes.POINTER(ctypes.c_double)
    return float(lib().pp_oracle_iou(a.ctypes.data_as(dp), g.ctypes.data_as(dp)))


def make_ious(a_corners, g_corners, a_centers, g_centers, ious):
    """Same signature and in-place semantics as ``pillars.make_ious`` (data/pillars.cpp:400-427)."""
    a_corners = np.ascontiguousarray(a_corners, dtype=np.float64)
    g_corners = np.ascontiguousarray(g_corners, dtype=np.float64)
    a_centers = np.ascontiguousarray(a_centers, dtype=np.float64)
    g_centers = np.ascontiguousarray(g_centers, dtype=np.float64)
    assert ious.dtype == np.float64 and ious.flags.c_contiguous
    assert ious.shape == (a_corners.shape[0], g_corners.shape[0])
    if g_corners.shape[0] == 0 or a_corners.shape[0] == 0:
        return
    rc = lib().pp_oracle_make_ious(a_corners.ctypes.data, g_corners.ctypes.data,
                                   a_centers.ctypes.data, g_centers.ctypes.data, ious.ctypes.data,
                                   a_corners.shape[0], g_corners.shape[0])
    if rc == 2:
        raise RuntimeError("IOU < 0 (reference would exit(1): data/pillars.cpp:166-169)")
    if rc != 0:
        raise RuntimeError("pp_oracle_make_ious failed: %d" % rc)
