"""Device-native batch entry points of the hot path: the calls a training / serving loop makes.

``InputPath`` owns the batch-invariant state (grid config, data_mean, anchors + index, PFN
parameters) on one GPU and turns raw sweeps + ground-truth boxes into the network's inputs:

    pillarize(points, offsets)          -> x [B,9,P,N] f32, inds [B,P,3] i64, n_pillars [B]
    encode(x, inds)                     -> canvas [B,C,H,W] f32           (PFN + scatter, fused)
    targets(gt batch)                   -> cls [B,A,K] f32, reg [B,A,9] f32
    step_host(points_list, gt_list)     -> the same three, from HOST (pinned) buffers

Sweeps are independent, so multi-GPU use is one ``InputPath`` per process / GPU over a shard of
the batch with no collective (``shard_range``).
"""
import numpy as np
import torch

from . import _lib, _runtime
from .box_utils import AnchorSet, assign_targets, gt_to_image_space
from .config import PPConfig
from .model import PPFeatureScatter


def shard_range(n_items, rank, world_size):
    """Contiguous chunk of ``n_items`` owned by ``rank`` (nn.DataParallel's chunking, train.py:88-89)."""
    per = (n_items + world_size - 1) // world_size
    lo = min(rank * per, n_items)
    return lo, min(lo + per, n_items)


class InputPath:
    def __init__(self, cfg=None, device=None, data_mean=None, pfn_params=None, anchors=None,
                 training=True):
        if not torch.cuda.is_available():
            raise _lib.PPError("no CUDA device: the input path has no CPU fallback")
        _lib.load()
        self.cfg = cfg or PPConfig()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        c = self.cfg
        self.data_mean = None
        if data_mean is not None:
            dm = torch.as_tensor(data_mean, dtype=torch.float32).reshape(-1)
            if dm.numel() != 9 * c.max_pillars * c.max_points_per_pillar:
                raise _lib.PPError("data_mean must have 9*P*N elements (make_means.py:28)")
            self.data_mean = dm.to(self.device)
        self.net = PPFeatureScatter(c.feature_net_in, c.feature_net_out, c.canvas_height,
                                    c.canvas_width).to(self.device)
        if pfn_params is not None:
            self.load_pfn_params(pfn_params)
        self.net.train(training)
        self.anchors = anchors
        self._pinned = {}
        # target assignment (K3) does not depend on K1/K2: it runs on a side stream and overlaps the
        # latency-bound binning kernels
        self._side = torch.cuda.Stream(device=self.device)
        self.overlap_targets = True

    # -- parameters ---------------------------------------------------------------------------
    def load_pfn_params(self, p):
        t = lambda a: torch.as_tensor(a, dtype=torch.float32, device=self.device)
        with torch.no_grad():
            self.net.conv1.weight.copy_(t(p["conv_w"]).reshape(self.net.conv1.weight.shape))
            self.net.conv1.bias.copy_(t(p["conv_b"]))
            self.net.bn1.weight.copy_(t(p["bn_w"]))
            self.net.bn1.bias.copy_(t(p["bn_b"]))
            self.net.bn1.running_mean.copy_(t(p["running_mean"]))
            self.net.bn1.running_var.copy_(t(p["running_var"]))

    def ensure_anchors(self):
        if self.anchors is None:
            self.anchors = AnchorSet.from_config(self.cfg, self.device)
        return self.anchors

    # -- K1 -----------------------------------------------------------------------------------
    def pillarize(self, points, offsets, out=None):
        """points: CUDA float32 [T, S] (S >= 4 columns: x,y,z,r,...) or float64, all sweeps
        concatenated; offsets: host list of len B+1.  Returns (x, inds, n_pillars)."""
        L = _lib.load()
        c = self.cfg
        _runtime.require_cuda(points, "points")
        if points.dim() != 2 or points.shape[1] < 4 or points.stride(1) != 1 and points.stride(0) != 1:
            raise _lib.PPError("points must be a 2-D [T, >=4] tensor")
        if points.dtype == torch.float32:
            dt = _lib.PP_F32
        elif points.dtype == torch.float64:
            dt = _lib.PP_F64
        else:
            raise _lib.PPError("points must be float32 or float64")
        B = len(offsets) - 1
        if B < 1 or B > _lib.PP_MAX_SWEEPS:
            raise _lib.PPError("1..%d sweeps per call" % _lib.PP_MAX_SWEEPS)
        T = int(offsets[-1])
        P, N = c.max_pillars, c.max_points_per_pillar
        dev = self.device
        if out is None:
            x = torch.empty((B, 9, P, N), dtype=torch.float32, device=dev)
            inds = torch.empty((B, P, 3), dtype=torch.int64, device=dev)
            npil = torch.empty(B, dtype=torch.int32, device=dev)
        else:
            x, inds, npil = out
        grid = c.grid()
        nbytes = L.pp_pillarize_workspace_bytes(B, T, grid, P)
        ws = _runtime.workspace(nbytes, dev, "pillarize")
        status = _runtime.status_word(dev)
        with torch.cuda.device(dev):
            rc = L.pp_pillarize(points.data_ptr() if T > 0 else None, dt, points.stride(0),
                                points.stride(1), _lib.i64_array(offsets), B, grid, N, P,
                                self.data_mean.data_ptr() if self.data_mean is not None else None,
                                x.data_ptr(), inds.data_ptr(), npil.data_ptr(), status.data_ptr(),
                                ws.data_ptr(), ws.numel(), _runtime.stream_ptr(dev))
        _lib.check(rc, "pp_pillarize")
        return x, inds, npil

    # -- K2 -----------------------------------------------------------------------------------
    def encode(self, x, inds, out=None, return_features=False):
        return self.net(x, inds, return_features=return_features, out=out)

    # -- K3 -----------------------------------------------------------------------------------
    def targets(self, gt_dev, gt_offsets, out=None):
        """gt_dev: dict of CUDA tensors corners [Gt,4,2], centers [Gt,3] (image space), wlh [Gt,3],
        yaw [Gt] (float64), cls [Gt] (int32)."""
        a = self.ensure_anchors()
        return assign_targets(a, gt_dev["corners"], gt_dev["centers"], gt_dev["wlh"], gt_dev["yaw"],
                              gt_dev["cls"], gt_offsets, self.cfg.num_classes,
                              self.cfg.iou_pos_thresh, out=out)

    # -- host-facing step -----------------------------------------------------------------------
    def pack_host_batch(self, sweeps, gts):
        """Pack a list of float32 [n_i, >=4] sweeps and a list of GT dicts (centers/wlh/yaw/cls in
        canvas space, as the reference's box pickles hold them) into pinned host buffers."""
        offs = [0]
        for s in sweeps:
            offs.append(offs[-1] + int(s.shape[0]))
        ncol = int(sweeps[0].shape[1])
        pts = torch.empty((max(offs[-1], 1), ncol), dtype=torch.float32).pin_memory()
        for s, lo, hi in zip(sweeps, offs[:-1], offs[1:]):
            pts[lo:hi] = torch.as_tensor(s, dtype=torch.float32)
        goffs = [0]
        for g in gts:
            goffs.append(goffs[-1] + int(len(g["yaw"])))
        Gt = goffs[-1]
        gpack = torch.zeros((max(Gt, 1), 16), dtype=torch.float64).pin_memory()   # corners 8 | centers 3 | wlh 3 | yaw 1 | cls 1
        for g, lo, hi in zip(gts, goffs[:-1], goffs[1:]):
            if hi == lo:
                continue
            cen, cor = gt_to_image_space(g, self.cfg.canvas_height)
            gpack[lo:hi, 0:8] = torch.from_numpy(cor.reshape(-1, 8))
            gpack[lo:hi, 8:11] = torch.from_numpy(cen)
            gpack[lo:hi, 11:14] = torch.from_numpy(np.asarray(g["wlh"], dtype=np.float64))
            gpack[lo:hi, 14] = torch.from_numpy(np.asarray(g["yaw"], dtype=np.float64))
            gpack[lo:hi, 15] = torch.from_numpy(np.asarray(g["cls"], dtype=np.float64))
        return {"points": pts, "offsets": offs, "gt": gpack, "gt_offsets": goffs}

    def _run(self, d_pts, offsets, gt_dev, gt_offsets, o):
        main = torch.cuda.current_stream(self.device)
        if self.overlap_targets:
            self.ensure_anchors()
            self._side.wait_stream(main)
            with torch.cuda.stream(self._side):
                cls, reg, top, counts = self.targets(gt_dev, gt_offsets, out=o.get("targets"))
            x, inds, npil = self.pillarize(d_pts, offsets, out=o.get("pillars"))
            canvas = self.encode(x, inds, out=o.get("canvas"))
            main.wait_stream(self._side)
            for t in (cls, reg, top, counts):
                t.record_stream(main)
        else:
            x, inds, npil = self.pillarize(d_pts, offsets, out=o.get("pillars"))
            canvas = self.encode(x, inds, out=o.get("canvas"))
            cls, reg, top, counts = self.targets(gt_dev, gt_offsets, out=o.get("targets"))
        return canvas, cls, reg, npil, counts

    def step_host(self, batch, out=None):
        """One pass of the whole path from pinned HOST buffers (``pack_host_batch``): H2D copy,
        pillarize, PFN + scatter, target assignment.  Outputs stay on the device, where the
        backbone and the loss consume them; returns (canvas, cls, reg, n_pillars, counts)."""
        dev = self.device
        T, Gt = batch["offsets"][-1], batch["gt_offsets"][-1]
        d_pts = batch["points"][:max(T, 1)].to(dev, non_blocking=True)
        d_gt = batch["gt"][:max(Gt, 1)].to(dev, non_blocking=True)
        gt_dev = {
            "corners": d_gt[:, 0:8].contiguous(), "centers": d_gt[:, 8:11].contiguous(),
            "wlh": d_gt[:, 11:14].contiguous(), "yaw": d_gt[:, 14].contiguous(),
            "cls": d_gt[:, 15].to(torch.int32),
        }
        return self._run(d_pts, batch["offsets"], gt_dev, batch["gt_offsets"], out or {})

    def step_device(self, d_pts, offsets, gt_dev, gt_offsets, out=None):
        """Same pass with inputs already resident in HBM."""
        return self._run(d_pts, offsets, gt_dev, gt_offsets, out or {})
