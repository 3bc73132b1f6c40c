"""ctypes binding of libpp_b200.so (include/pp_b200.h).

There is no CPU fallback: if the shared object is missing and cannot be built, or a call returns
an error, this module raises.  PyTorch is only used by the callers for device memory and streams.
"""
import ctypes
import os

from . import build as _build

PP_OK = 0
PP_F32, PP_F64, PP_I64 = 0, 1, 2
PP_MAX_SWEEPS = 64
STATUS_NEG_IOU, STATUS_BAD_POINT, STATUS_BAD_INDEX, STATUS_CAND_OVERFLOW, STATUS_RANGE = 1, 2, 4, 8, 16


class PPGrid(ctypes.Structure):
    """pp_grid: the nine doubles create_pillars takes positionally (data/pillars.cpp:241-249)."""
    _fields_ = [(n, ctypes.c_double) for n in (
        "x_step", "y_step", "x_min", "y_min", "z_min", "x_max", "y_max", "z_max", "canvas_height")]


class PPError(RuntimeError):
    pass


_vp0, _i320, _i640, _sz0, _f320, _f640 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_size_t, ctypes.c_float, ctypes.c_double


class PPStepPlan(ctypes.Structure):
    """pp_step_plan (include/pp_b200.h), field for field."""
    _fields_ = [
        ("stream_main", _vp0), ("stream_side", _vp0), ("stream_copy", _vp0),
        ("ev_fork", _vp0), ("ev_join", _vp0), ("ev_ready", _vp0), ("ev_slot_free", _vp0), ("ev_prev_encode", _vp0),
        ("ev_encode_done", _vp0), ("ev_done", _vp0),
        ("h_blob", _vp0), ("d_blob", _vp0), ("blob_bytes", _sz0),
        ("n_files", _i320), ("d_file_offsets", _vp0), ("d_file_xforms", _vp0), ("min_dist", _f320),
        ("d_points", _vp0), ("total_points", _i640), ("point_cols", _i640), ("h_sweep_offsets", _vp0), ("n_sweeps", _i320),
        ("grid", PPGrid), ("max_points_per_pillar", _i320), ("max_pillars", _i320), ("d_data_mean", _vp0),
        ("d_mean_prepared", _vp0),
        ("C", _i320), ("d_conv_w", _vp0), ("d_conv_b", _vp0), ("d_bn_w", _vp0), ("d_bn_b", _vp0), ("d_running_mean", _vp0),
        ("d_running_var", _vp0),
        ("d_num_batches_tracked", _vp0), ("training", _i320), ("momentum", _f320), ("eps", _f320), ("canvas_h", _i320),
        ("canvas_w", _i320),
        ("d_canvas", _vp0), ("d_indices", _vp0), ("d_num_pillars", _vp0), ("d_ws_input", _vp0), ("ws_input_bytes", _sz0),
        ("d_a_corners", _vp0), ("d_a_centers", _vp0), ("d_a_wlh", _vp0), ("d_a_yaw", _vp0), ("d_anchor_index", _vp0), ("A", _i640),
        ("d_g_corners", _vp0), ("d_g_centers", _vp0), ("d_g_wlh", _vp0), ("d_g_yaw", _vp0), ("d_g_cls", _vp0),
        ("h_gt_offsets", _vp0),
        ("num_classes", _i320), ("pos_thresh", _f640), ("d_cls", _vp0), ("d_reg", _vp0), ("d_top_anchor", _vp0), ("d_counts", _vp0),
        ("d_pos_anchor", _vp0), ("d_pos_cls", _vp0), ("d_pos_reg", _vp0), ("d_pos_offsets", _vp0), ("pos_capacity", _i320),
        ("d_ws_targets", _vp0), ("ws_targets_bytes", _sz0),
        ("d_status", _vp0), ("h_counters", _vp0)]


_lib = None
_c = ctypes
_vp, _i32, _i64, _sz, _f32, _f64 = _c.c_void_p, _c.c_int32, _c.c_int64, _c.c_size_t, _c.c_float, _c.c_double
_gridp = _c.POINTER(PPGrid)
_i64p = _c.POINTER(_c.c_int64)

_SIGNATURES = {
    "pp_version": (_c.c_int, []),
    "pp_error_string": (_c.c_char_p, [_c.c_int]),
    "pp_last_cuda_error": (_c.c_int, []),
    "pp_launch_count": (_i64, []),
    "pp_set_option": (_c.c_int, [_c.c_char_p, _c.c_int]),
    "pp_profile_enable": (_c.c_int, [_c.c_int]),
    "pp_profile_report": (_i64, [_c.c_char_p, _i64]),
    "pp_pillarize_workspace_bytes": (_sz, [_i32, _i64, _gridp, _i32]),
    "pp_pillarize": (_c.c_int, [_vp, _i32, _i64, _i64, _i64p, _i32, _gridp, _i32, _i32, _vp, _vp,
                                _vp, _vp, _vp, _vp, _sz, _vp]),
    "pp_pillarize_compact": (_c.c_int, [_vp, _i32, _i64, _i64, _i64, _gridp, _i32, _i32, _vp, _vp,
                                        _vp, _vp, _vp, _vp, _sz, _vp]),
    "pp_pfn_workspace_bytes": (_sz, [_i32, _i32, _i32, _i32, _i32]),
    "pp_pfn_forward": (_c.c_int, [_vp, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp,
                                  _vp, _i32, _f32, _f32, _vp, _vp, _sz, _vp]),
    "pp_scatter": (_c.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _sz, _vp]),
    "pp_pfn_scatter": (_c.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp,
                                  _vp, _vp, _i32, _f32, _f32, _i32, _i32, _vp, _vp, _vp, _vp, _sz,
                                  _vp]),
    "pp_input_path_workspace_bytes": (_sz, [_i32, _i64, _gridp, _i32, _i32, _i32, _i32, _i32, _i32]),
    "pp_mean_prepared_bytes": (_sz, [_i32, _i32]),
    "pp_mean_prepare": (_c.c_int, [_vp, _i32, _i32, _vp, _sz, _vp]),
    "pp_input_path": (_c.c_int, [_vp, _i32, _i64, _i64, _i64p, _i32, _gridp, _i32, _i32, _vp, _vp, _i32,
                                 _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _f32, _f32, _i32, _i32,
                                 _vp, _vp, _vp, _vp, _vp, _vp, _sz, _i32, _vp]),
    "pp_step": (_c.c_int, [_c.POINTER(PPStepPlan)]),
    "pp_step_plan_bytes": (_sz, []),
    "pp_input_path_backward_workspace_bytes": (_sz, [_i32, _i32, _i32]),
    "pp_input_path_backward": (_c.c_int, [_i64p, _i32, _gridp, _i32, _i32, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _i32, _f32, _i32,
                                          _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp, _sz, _vp]),
    "pp_aggregate_sweeps": (_c.c_int, [_vp, _i64, _i32, _vp, _i32, _vp, _f32, _vp, _vp]),
    "pp_pfn_backward_workspace_bytes": (_sz, [_i32, _i32, _i32]),
    "pp_pfn_backward": (_c.c_int, [_vp, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _i32, _f32, _vp, _vp,
                                   _i32, _i32, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "pp_scatter_backward": (_c.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp, _vp]),
    "pp_loss_workspace_bytes": (_sz, [_i32, _i32, _i32, _i32]),
    "pp_loss_scale_grads": (_c.c_int, [_vp, _sz, _vp, _sz, _vp, _vp, _vp]),
    "pp_loss": (_c.c_int, [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _f32, _f32, _f32, _f32, _f32,
                           _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "pp_loss_list": (_c.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _f32, _f32, _f32, _f32,
                                _f32, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "pp_make_ious": (_c.c_int, [_vp, _vp, _vp, _vp, _i64, _i64, _vp, _vp, _vp]),
    "pp_anchor_index_bytes": (_sz, [_vp, _i64, _i32]),
    "pp_anchor_index_build": (_c.c_int, [_vp, _vp, _i64, _vp, _sz, _vp]),
    "pp_assign_targets_workspace_bytes": (_sz, [_i32, _i64, _i64, _vp]),
    "pp_assign_targets": (_c.c_int, [_vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _i64p,
                                     _i32, _i32, _f64, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "pp_assign_targets_list": (_c.c_int, [_vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _i64p,
                                          _i32, _i32, _f64, _vp, _vp, _vp, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _sz,
                                          _vp]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)

# include/pp_b200_debug.h: bound only when the library was built with -DPP_DEBUG (build.py --debug)
_DEBUG_SIGNATURES = {
    "pp_debug_set": (_c.c_int, [_c.c_char_p, _c.c_int]),
    "pp_debug_tc_timing": (_c.c_int, [_i64p]),
}


def lib_path():
    return _build.LIB_PATH


def load():
    """Load libpp_b200.so, building it with nvcc first if it is missing or stale."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB_PATH
    if not os.path.exists(path) or (_build.is_stale() and os.access(os.path.dirname(path), os.W_OK)):
        try:
            _build.build()
        except Exception as e:  # noqa: BLE001
            if not os.path.exists(path):
                raise PPError("libpp_b200.so is missing and could not be built (no CPU fallback "
                              "exists): %s" % e) from e
    L = ctypes.CDLL(path)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(L, name)  # AttributeError if a declared symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if L.pp_step_plan_bytes() != ctypes.sizeof(PPStepPlan):
        raise PPError("pp_step_plan layout mismatch: C %d bytes, ctypes %d" % (L.pp_step_plan_bytes(), ctypes.sizeof(PPStepPlan)))
    for name, (res, args) in _DEBUG_SIGNATURES.items():
        fn = getattr(L, name, None)
        if fn is not None:
            fn.restype = res
            fn.argtypes = args
    _lib = L
    return L


def check(rc, what):
    if rc != PP_OK:
        L = load()
        msg = L.pp_error_string(rc).decode()
        raise PPError("%s failed: %s (code %d, cudaError %d)" % (what, msg, rc, L.pp_last_cuda_error()))


_i64_cache = {}


def i64_array(values):
    """ctypes int64 array of a short host sequence (sweep / GT offsets); the same offsets recur step after step."""
    key = tuple(int(v) for v in values)
    arr = _i64_cache.get(key)
    if arr is None:
        if len(_i64_cache) > 256:
            _i64_cache.clear()
        arr = _i64_cache[key] = (ctypes.c_int64 * len(key))(*key)
    return arr


def profile_report():
    """{kernel name: (launches, total_ms)} since pp_profile_enable(1); synchronises the device."""
    L = load()
    buf = ctypes.create_string_buffer(1 << 16)
    L.pp_profile_report(buf, len(buf))
    out = {}
    for line in buf.value.decode().splitlines():
        name, n, ms = line.split()
        out[name] = (int(n), float(ms))
    return out
