"""Host mirror of the reference's target assignment (utils/box_utils.py:19-32,70-109,111-159,
162-232) over the sm_100a kernels of libpp_b200.so.

``create_target`` keeps the reference's signature (numpy arrays + duck-typed box lists exposing
``center``, ``wlh``, ``name``, ``orientation.yaw_pitch_roll[0]``, ``bottom_corners()``);
``AnchorSet`` / ``assign_targets`` are the device-native batch entry points.  No CPU fallback.
"""
import numpy as np
import torch

from . import _lib, _runtime
from .config import cfg as _cfg


def make_anchor_arrays(cfg=None):
    """The anchor lattice of utils/box_utils.py:111-159 as arrays (no Box objects):
    anchor a = (y*fm_width + x)*n_d + d, centre ((x+.5)/fm_scale, (y+.5)/fm_scale, z_d), corners in
    the bottom_corners() order (+l/2,-w/2), (+l/2,+w/2), (-l/2,+w/2), (-l/2,-w/2) rotated by yaw.
    Returns dict: corners [A,4,2], centers [A,3], wlh [A,3], yaw [A] (radians), all float64."""
    cfg = cfg or _cfg
    nd = len(cfg.anchor_dims)
    ys, xs, ds = np.meshgrid(np.arange(cfg.fm_height), np.arange(cfg.fm_width), np.arange(nd),
                             indexing="ij")
    ys, xs, ds = ys.reshape(-1), xs.reshape(-1), ds.reshape(-1)
    dims = np.stack([np.asarray(d, dtype=np.float64) for d in cfg.anchor_dims])
    wlh = dims[ds]
    yaw = np.array([_sdk_yaw(float(d)) for d in cfg.anchor_yaws_deg], dtype=np.float64)[ds]
    centers = np.stack([(xs + 0.5) / cfg.fm_scale, (ys + 0.5) / cfg.fm_scale,
                        np.asarray(cfg.anchor_zs, dtype=np.float64)[ds]], axis=1)
    corners = box_corners(centers, wlh, yaw)
    return {"corners": corners, "centers": centers, "wlh": wlh, "yaw": yaw}


def _sdk_yaw(deg):
    """The yaw the reference reads back from an anchor built as ``Quaternion(axis=[0,0,1], degrees=deg)``
    (utils/box_utils.py:80,147): pyquaternion's ``yaw_pitch_roll[0] = atan2(2 w z, 1 - 2 z^2)`` of
    ``(w, z) = (cos(a/2), sin(a/2))``, ``a = deg/180*pi``.  For 90 degrees that is pi/2 - 2.2e-16, not pi/2."""
    a = deg / 180.0 * np.pi
    w, z = np.cos(a / 2.0), np.sin(a / 2.0)
    return float(np.arctan2(2 * (w * z), 1 - 2 * (z ** 2)))


def box_corners(centers, wlh, yaw):
    """Bottom corners [n,4,2] in the Lyft Box.bottom_corners() order, counter-clockwise."""
    w, l = wlh[:, 0:1], wlh[:, 1:2]
    xs = l / 2 * np.array([[1, 1, -1, -1.0]])
    ys = w / 2 * np.array([[-1, 1, 1, -1.0]])
    c, s = np.cos(yaw)[:, None], np.sin(yaw)[:, None]
    return np.stack([c * xs - s * ys + centers[:, 0:1], s * xs + c * ys + centers[:, 1:2]], axis=2)


def boxes_to_image_space(boxes, canvas_height=None):
    """utils/box_utils.py:19-32."""
    H = _cfg.canvas_height if canvas_height is None else canvas_height
    centers = np.stack([np.asarray(box.center, dtype=np.float64).copy() for box in boxes])
    corners = np.stack([box.bottom_corners().transpose([1, 0])[:, :2] for box in boxes])
    centers[..., 1] = (H - 1) - centers[..., 1]
    corners[..., 1] = (H - 1) - corners[..., 1]
    return centers, corners


def gt_to_image_space(gt, canvas_height=None):
    """Array form of boxes_to_image_space for a dict(centers, wlh, yaw, cls): returns
    (centers_img [G,3], corners_img [G,4,2]); the y flip turns the CCW ring clockwise."""
    H = _cfg.canvas_height if canvas_height is None else canvas_height
    centers = np.array(gt["centers"], dtype=np.float64, copy=True).reshape(-1, 3)
    corners = box_corners(centers, np.asarray(gt["wlh"], dtype=np.float64).reshape(-1, 3),
                          np.asarray(gt["yaw"], dtype=np.float64).reshape(-1))
    centers[:, 1] = (H - 1) - centers[:, 1]
    corners[..., 1] = (H - 1) - corners[..., 1]
    return centers, corners


class AnchorSet:
    """Device-resident anchors + the bucket index the assignment kernels walk.  Built once per
    anchor set (the reference builds and pickles its anchors once, train_prep.py:115-120)."""

    def __init__(self, corners, centers, wlh, yaw, device=None):
        L = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.PPError("no CUDA device: target assignment has no CPU fallback")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.h_centers = np.ascontiguousarray(centers, dtype=np.float64)
        self.A = int(self.h_centers.shape[0])
        f = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(self.device)
        self.corners, self.centers, self.wlh, self.yaw = f(corners), f(self.h_centers), f(wlh), f(yaw)
        h_corners = np.ascontiguousarray(corners, dtype=np.float64).reshape(self.A, 8)
        nbytes = L.pp_anchor_index_bytes(self.h_centers.ctypes.data, self.A, 1)
        if nbytes == 0:
            raise _lib.PPError("pp_anchor_index_bytes: invalid anchors")
        self.index = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            rc = L.pp_anchor_index_build(self.h_centers.ctypes.data, h_corners.ctypes.data, self.A,
                                         self.index.data_ptr(), self.index.numel(), _runtime.stream_ptr(self.device))
        _lib.check(rc, "pp_anchor_index_build")

    @classmethod
    def from_config(cls, cfg=None, device=None):
        a = make_anchor_arrays(cfg)
        return cls(a["corners"], a["centers"], a["wlh"], a["yaw"], device=device)


class Positives(object):
    """The targets as a positives list + implicit zeros (pp_assign_targets_list): ``anchor [cap]`` int32 =
    sweep * A + anchor in ascending order, ``cls [cap,9]`` / ``reg [cap,9]`` float32 rows, ``offsets [B+1]`` int32
    (device; ``offsets[B]`` is the length), plus the shape ``(B, A)`` they stand for.  ``pp_b200.loss.PPLoss``
    takes it in place of the two dense tensors."""

    def __init__(self, anchor, cls, reg, offsets, n_sweeps, n_anchors):
        self.anchor, self.cls, self.reg, self.offsets = anchor, cls, reg, offsets
        self.n_sweeps, self.n_anchors = int(n_sweeps), int(n_anchors)

    def dense(self):
        """(cls [B,A,9], reg [B,A,9]) float32: what pp_assign_targets would have written (synchronises; raises if
        the list overflowed its capacity -- offsets are clamped to it, the dropped rows are lost)."""
        n = int(self.offsets[-1].item())
        if self.anchor.is_cuda:
            _runtime.check_status(self.anchor.device, "Positives.dense")
        B, A = self.n_sweeps, self.n_anchors
        cls = torch.zeros((B * A, 9), dtype=torch.float32, device=self.anchor.device)
        reg = torch.zeros((B * A, 9), dtype=torch.float32, device=self.anchor.device)
        idx = self.anchor[:n].long()
        cls[idx] = self.cls[:n]
        reg[idx] = self.reg[:n]
        return cls.view(B, A, 9), reg.view(B, A, 9)


def assign_targets(anchors, g_corners, g_centers, g_wlh, g_yaw, g_cls, gt_offsets, num_classes=None,
                   pos_thresh=None, out=None, as_list=False, capacity=None):
    """Device-native batch target assignment (pp_assign_targets).  GT tensors are CUDA float64
    (g_cls int32), all sweeps concatenated; ``gt_offsets`` is a host list of len n_sweeps+1.
    Returns (cls [B,A,K] f32, reg [B,A,9] f32, top_anchor [Gt] i32, counts [B,4] i32).
    ``as_list=True`` (pp_assign_targets_list): returns (Positives, None, top_anchor, counts) and writes no
    dense tensor; ``capacity`` bounds the list (default 4096 per sweep)."""
    if as_list:
        return _assign_targets_list(anchors, g_corners, g_centers, g_wlh, g_yaw, g_cls, gt_offsets, num_classes,
                                    pos_thresh, capacity)
    L = _lib.load()
    dev = anchors.device
    B = len(gt_offsets) - 1
    K = int(_cfg.num_classes if num_classes is None else num_classes)
    thr = float(_cfg.iou_pos_thresh if pos_thresh is None else pos_thresh)
    Gt = int(gt_offsets[-1])
    A = anchors.A
    if out is None:
        cls = torch.empty((B, A, K), dtype=torch.float32, device=dev)
        reg = torch.empty((B, A, 9), dtype=torch.float32, device=dev)
    else:
        cls, reg = out
    top = torch.empty(max(Gt, 1), dtype=torch.int32, device=dev)
    counts = torch.empty((B, 4), dtype=torch.int32, device=dev)
    nbytes = L.pp_assign_targets_workspace_bytes(B, A, Gt, None)
    ws = _runtime.workspace(nbytes, dev, "targets")
    status = _runtime.status_word(dev)
    ptr = lambda t: t.data_ptr() if (t is not None and t.numel() > 0) else None
    with _runtime.on_device(dev):
        rc = L.pp_assign_targets(
            anchors.corners.data_ptr(), anchors.centers.data_ptr(), anchors.wlh.data_ptr(),
            anchors.yaw.data_ptr(), anchors.index.data_ptr(), A, ptr(g_corners), ptr(g_centers),
            ptr(g_wlh), ptr(g_yaw), ptr(g_cls), _lib.i64_array(gt_offsets), B, K, thr,
            cls.data_ptr(), reg.data_ptr(), top.data_ptr(), counts.data_ptr(), status.data_ptr(),
            ws.data_ptr(), ws.numel(), _runtime.stream_ptr(dev))
    _lib.check(rc, "pp_assign_targets")
    _runtime.poll_status(dev, "assign_targets")
    return cls, reg, top[:Gt], counts


def _assign_targets_list(anchors, g_corners, g_centers, g_wlh, g_yaw, g_cls, gt_offsets, num_classes, pos_thresh,
                         capacity):
    L = _lib.load()
    dev = anchors.device
    B = len(gt_offsets) - 1
    K = int(_cfg.num_classes if num_classes is None else num_classes)
    thr = float(_cfg.iou_pos_thresh if pos_thresh is None else pos_thresh)
    Gt = int(gt_offsets[-1])
    A = anchors.A
    cap = int(capacity if capacity is not None else 4096 * B)
    pos = Positives(torch.empty(cap, dtype=torch.int32, device=dev), torch.empty((cap, 9), dtype=torch.float32, device=dev),
                    torch.empty((cap, 9), dtype=torch.float32, device=dev),
                    torch.empty(B + 1, dtype=torch.int32, device=dev), B, A)
    top = torch.empty(max(Gt, 1), dtype=torch.int32, device=dev)
    counts = torch.empty((B, 4), dtype=torch.int32, device=dev)
    ws = _runtime.workspace(L.pp_assign_targets_workspace_bytes(B, A, Gt, None), dev, "targets")
    status = _runtime.status_word(dev)
    ptr = lambda t: t.data_ptr() if (t is not None and t.numel() > 0) else None
    with _runtime.on_device(dev):
        rc = L.pp_assign_targets_list(
            anchors.corners.data_ptr(), anchors.centers.data_ptr(), anchors.wlh.data_ptr(),
            anchors.yaw.data_ptr(), anchors.index.data_ptr(), A, ptr(g_corners), ptr(g_centers),
            ptr(g_wlh), ptr(g_yaw), ptr(g_cls), _lib.i64_array(gt_offsets), B, K, thr,
            pos.anchor.data_ptr(), pos.cls.data_ptr(), pos.reg.data_ptr(), pos.offsets.data_ptr(), cap, None, None,
            top.data_ptr(), counts.data_ptr(), status.data_ptr(), ws.data_ptr(), ws.numel(), _runtime.stream_ptr(dev))
    _lib.check(rc, "pp_assign_targets_list")
    _runtime.poll_status(dev, "assign_targets(as_list=True): positives list capacity %d" % cap)
    return pos, None, top[:Gt], counts


_anchor_cache = {}


def _anchor_set_for(anchor_corners, anchor_centers, anchor_box_list):
    if isinstance(anchor_box_list, AnchorSet):
        return anchor_box_list
    key = (id(anchor_box_list), id(anchor_corners), len(anchor_box_list))
    hit = _anchor_cache.get(key)
    if hit is None:
        wlh = np.stack([np.asarray(b.wlh, dtype=np.float64) for b in anchor_box_list])
        yaw = np.array([b.orientation.yaw_pitch_roll[0] for b in anchor_box_list], dtype=np.float64)
        hit = AnchorSet(anchor_corners, anchor_centers, wlh, yaw)
        _anchor_cache.clear()
        _anchor_cache[key] = hit
    return hit


def create_target(anchor_corners, gt_corners, anchor_centers, gt_centers, anchor_box_list,
                  gt_box_list):
    """utils/box_utils.py:162-232, same arguments and return ``(cls_targets [A,K], reg_targets
    [A,9])`` as float64 numpy arrays holding the float32 values the reference's caller keeps
    (data/dataset.py:117-118).  ``anchor_box_list`` may be an ``AnchorSet`` (the list form is
    converted once and cached)."""
    anchors = _anchor_set_for(anchor_corners, anchor_centers, anchor_box_list)
    dev = anchors.device
    G = len(gt_box_list)
    f = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(dev)
    if G > 0:
        name_to_ind = _cfg.name_to_ind
        g_cls = torch.tensor([name_to_ind[b.name] for b in gt_box_list], dtype=torch.int32, device=dev)
        g_wlh = f(np.stack([np.asarray(b.wlh, dtype=np.float64) for b in gt_box_list]))
        g_yaw = f(np.array([b.orientation.yaw_pitch_roll[0] for b in gt_box_list]))
        g_cor, g_cen = f(np.asarray(gt_corners).reshape(G, 4, 2)), f(np.asarray(gt_centers).reshape(G, 3))
    else:
        g_cls = g_wlh = g_yaw = g_cor = g_cen = None
    cls, reg, _, _ = assign_targets(anchors, g_cor, g_cen, g_wlh, g_yaw, g_cls, [0, G])
    out = cls[0].cpu().numpy().astype(np.float64), reg[0].cpu().numpy().astype(np.float64)
    _runtime.check_status(dev, "create_target")
    return out
