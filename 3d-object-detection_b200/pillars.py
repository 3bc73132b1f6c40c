"""Drop-in for the reference's pybind11 module ``data.pillars`` (data/pillars.cpp:429-435).

``create_pillars`` and ``make_ious`` keep the reference's positional signatures and in-place numpy
semantics (caller allocates float64 outputs, callee mutates them), but run on the GPU: host
arrays are copied to the device, the sm_100a kernels of libpp_b200.so run, results are copied
back.  These two functions exist for signature compatibility and parity testing; the fast path
keeps data on the device (``pipeline.InputPath``).  There is no CPU fallback.
"""
import warnings

import numpy as np
import torch

from . import _lib, _runtime


def _device():
    if not torch.cuda.is_available():
        raise _lib.PPError("no CUDA device: the pillars drop-in has no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _writable_f64(arr, name):
    """pybind11 ``py::array_t<double>&`` with default forcecast: a non-float64 array is silently
    copied and the writes are lost.  We warn instead of being silent."""
    if not isinstance(arr, np.ndarray) or arr.dtype != np.float64:
        warnings.warn("%s is not a float64 numpy array: the reference would write into a "
                      "temporary copy and the result would be lost; nothing written" % name)
        return False
    return True


def create_pillars(points, tensor, indices, max_points_per_pillar, max_pillars, x_step, y_step,
                   x_min, y_min, z_min, x_max, y_max, z_max, canvas_height):
    """data/pillars.cpp:236-398.  ``points`` float64 [Npts, >=4] (any strides); ``tensor``
    float64 [>=max_pillars, >=max_points_per_pillar, 9] and ``indices`` float64 [>=max_pillars, 3]
    are mutated in place; only touched slots are written.  Pillar order is first-touch order
    (DESIGN.md), the reference's being Boost-hash order."""
    dev = _device()
    L = _lib.load()
    pts = np.asarray(points, dtype=np.float64)
    if pts.ndim != 2 or pts.shape[1] < 4:
        raise IndexError("points must be [Npts, >=4]")   # pybind11 .at() would raise index_error
    ok_t = _writable_f64(tensor, "tensor")
    ok_i = _writable_f64(indices, "indices")
    N, P = int(max_points_per_pillar), int(max_pillars)
    if ok_t and (tensor.ndim != 3 or tensor.shape[2] < 9):
        raise IndexError("tensor must be [P, N, 9]")
    if ok_i and (indices.ndim != 2 or indices.shape[1] < 3):
        raise IndexError("indices must be [P, 3]")
    n = pts.shape[0]
    if N < 1 or P < 1:
        return
    grid = _lib.PPGrid(float(x_step), float(y_step), float(x_min), float(y_min), float(z_min),
                       float(x_max), float(y_max), float(z_max), float(canvas_height))
    d_pts = torch.from_numpy(np.ascontiguousarray(pts[:, :4])).to(dev)
    d_rows = torch.empty((max(n, 1), 9), dtype=torch.float64, device=dev)
    d_slot = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
    d_xy = torch.empty((P, 2), dtype=torch.int32, device=dev)
    d_counts = torch.zeros(2, dtype=torch.int32, device=dev)
    status = _runtime.status_word(dev)
    nbytes = L.pp_pillarize_workspace_bytes(1, n, grid, P)
    if nbytes == 0:
        raise _lib.PPError("create_pillars: invalid grid / sizes")
    ws = _runtime.workspace(nbytes, dev, "pillarize")
    rc = L.pp_pillarize_compact(d_pts.data_ptr(), _lib.PP_F64, 4, 1, n, grid, N, P,
                                d_rows.data_ptr(), d_slot.data_ptr(), d_xy.data_ptr(),
                                d_counts.data_ptr(), status.data_ptr(), ws.data_ptr(), ws.numel(),
                                _runtime.stream_ptr(dev))
    _lib.check(rc, "pp_pillarize_compact")
    counts = d_counts.cpu().numpy()
    _runtime.check_status(dev, "create_pillars")
    n_pillars = int(counts[0])
    if ok_t and n > 0:
        slot = d_slot[:n].cpu().numpy()
        valid = slot >= 0
        if valid.any():
            rows = d_rows[:n].cpu().numpy()[valid]
            s = slot[valid].astype(np.int64)
            if (s // N).max() >= tensor.shape[0] or N > tensor.shape[1]:
                raise IndexError("tensor too small for max_pillars / max_points_per_pillar")
            tensor[s // N, s % N, :9] = rows
    if ok_i and n_pillars > 0:
        if n_pillars > indices.shape[0]:
            raise IndexError("indices too small for max_pillars")
        xy = d_xy[:n_pillars].cpu().numpy()
        indices[:n_pillars, 0] = 1
        indices[:n_pillars, 1] = xy[:, 0]
        indices[:n_pillars, 2] = xy[:, 1]


def make_ious(a_corners, g_corners, a_centers, g_centers, ious):
    """data/pillars.cpp:400-427.  Fills every entry of the caller's float64 ``ious`` [A,G]."""
    dev = _device()
    L = _lib.load()
    a_corners = np.ascontiguousarray(a_corners, dtype=np.float64)
    g_corners = np.ascontiguousarray(g_corners, dtype=np.float64)
    a_centers = np.ascontiguousarray(a_centers, dtype=np.float64)
    g_centers = np.ascontiguousarray(g_centers, dtype=np.float64)
    A, G = a_corners.shape[0], g_corners.shape[0]
    if a_corners.shape[1:] != (4, 2) or (G > 0 and g_corners.shape[1:] != (4, 2)):
        raise IndexError("corners must be [.,4,2]")
    if a_centers.shape[0] < A or g_centers.shape[0] < G or a_centers.shape[1] < 2:
        raise IndexError("centers must be [.,>=2]")
    if not _writable_f64(ious, "ious"):
        return
    if ious.ndim != 2 or ious.shape[0] < A or ious.shape[1] < G:
        raise IndexError("ious must be [A,G]")
    if A == 0 or G == 0:
        return
    if a_centers.shape[1] != 3:
        a_centers = np.ascontiguousarray(np.pad(a_centers[:, :2], ((0, 0), (0, 1))))
    if g_centers.shape[1] != 3:
        g_centers = np.ascontiguousarray(np.pad(g_centers[:, :2], ((0, 0), (0, 1))))
    d = [torch.from_numpy(x).to(dev) for x in (a_corners, g_corners, a_centers, g_centers)]
    d_out = torch.empty((A, G), dtype=torch.float64, device=dev)
    status = _runtime.status_word(dev)
    rc = L.pp_make_ious(d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), d[3].data_ptr(), A, G,
                        d_out.data_ptr(), status.data_ptr(), _runtime.stream_ptr(dev))
    _lib.check(rc, "pp_make_ious")
    ious[:A, :G] = d_out.cpu().numpy()
    _runtime.check_status(dev, "make_ious")
