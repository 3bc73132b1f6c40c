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
import ctypes

import numpy as np
import torch

from . import _lib, _runtime
from .box_utils import AnchorSet, assign_targets, gt_to_image_space
from .config import PPConfig
from .model import PPFeatureScatter, _bn_args


def shard_range(n_items, rank, world_size):
    """Contiguous chunk of ``n_items`` owned by ``rank`` (nn.DataParallel's chunking, train.py:88-89)."""
    per = (n_items + world_size - 1) // world_size
    lo = min(rank * per, n_items)
    return lo, min(lo + per, n_items)


class StepHandle:
    """Result of ``InputPath.step_host_async``: device outputs + the step's counters on their way to
    pinned host memory."""

    def __init__(self, outputs, pinned, n_sweeps, done):
        self.canvas, self.cls, self.reg, self.n_pillars, self.counts = outputs
        self._pin, self._B, self._done = pinned, n_sweeps, done

    def wait(self, stream=None):
        """Order ``stream`` (default: the current one) after this step; the outputs are then safe to use there.
        The outputs were allocated on a lane stream: they are recorded on ``stream`` so that the caching allocator
        does not hand their memory to a later step of that lane while kernels queued on ``stream`` still read it."""
        stream = stream or torch.cuda.current_stream()
        stream.wait_event(self._done)
        for t in self._tensors():
            t.record_stream(stream)

    def _tensors(self):
        out = []
        for t in (self.canvas, self.cls, self.reg, self.n_pillars, self.counts):
            if isinstance(t, torch.Tensor):
                out.append(t)
            elif t is not None and hasattr(t, "anchor"):                    # box_utils.Positives
                out += [t.anchor, t.cls, t.reg, t.offsets]
        return out

    def synchronize(self):
        self._done.synchronize()

    def counters(self):
        """(n_pillars [B], counts [B,4]) as host int32 tensors; waits for this step's D2H copy.  The device status
        word travels with them: a set bit (scatter index outside the canvas, NaN point, positives-list overflow,
        value outside the fp16 range of the padding pass ...) raises here."""
        self._done.synchronize()
        B = self._B
        if self._pin is None:
            _runtime.check_status(self.n_pillars.device, "input path step")
            return self.n_pillars.cpu(), self.counts.cpu()
        status = int(self._pin[5 * B])
        if status:
            _runtime.status_word(self.n_pillars.device).zero_()
            self._pin[5 * B] = 0
            _runtime._raise_status(status, "input path step")
        return self._pin[:B].clone(), self._pin[B:5 * B].view(B, 4).clone()


class _InputPathFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, conv_w, conv_b, bn_w, bn_b, path, points, offsets):
        net = path.net
        ctx.training = bool(net.training)
        ctx.eps = float(net.bn1.eps)
        ctx.stats = None if ctx.training else (net.bn1.running_mean.clone(), net.bn1.running_var.clone())
        canvas, inds, npil = path.pillarize_encode(points, offsets)
        ctx.ws = _runtime.take_workspace(path.device, "input_path")      # K1's compact state lives in it
        ctx.path, ctx.offsets = path, offsets
        ctx.save_for_backward(conv_w, conv_b, bn_w, inds, npil)
        ctx.mark_non_differentiable(inds, npil)
        return canvas, inds, npil

    @staticmethod
    def backward(ctx, g_canvas, _gi, _gn):
        L = _lib.load()
        conv_w, conv_b, bn_w, inds, npil = ctx.saved_tensors
        path, offsets = ctx.path, ctx.offsets
        c = path.cfg
        B = len(offsets) - 1
        P, N, C = c.max_pillars, c.max_points_per_pillar, c.feature_net_out
        dev = g_canvas.device
        g_canvas = g_canvas.contiguous()
        if g_canvas.dtype != torch.float32:
            raise _lib.PPError("input path backward: the canvas gradient must be float32")
        g_w = torch.empty((C, 9), dtype=torch.float32, device=dev)
        g_b, g_gamma, g_beta = (torch.empty(C, dtype=torch.float32, device=dev) for _ in range(3))
        ws2 = _runtime.workspace(L.pp_input_path_backward_workspace_bytes(B, P, C), dev, "input_path_bwd")
        rm, rv = ctx.stats if ctx.stats is not None else (None, None)
        w2 = conv_w.detach().reshape(C, 9).contiguous()
        with _runtime.on_device(dev):
            rc = L.pp_input_path_backward(
                _lib.i64_array(offsets), B, path._grid_struct(), N, P, path.data_mean.data_ptr() if path.data_mean is not None else None,
                C, w2.data_ptr(), conv_b.data_ptr(), bn_w.data_ptr(), rm.data_ptr() if rm is not None else None,
                rv.data_ptr() if rv is not None else None, 1 if ctx.training else 0, ctx.eps, c.canvas_height,
                c.canvas_width, g_canvas.data_ptr(), inds.data_ptr(), npil.data_ptr(), g_w.data_ptr(), g_b.data_ptr(),
                g_gamma.data_ptr(), g_beta.data_ptr(), ctx.ws.data_ptr(), ctx.ws.numel(), ws2.data_ptr(), ws2.numel(),
                _runtime.stream_ptr(dev))
        _lib.check(rc, "pp_input_path_backward")
        ctx.ws = None
        return g_w.view(C, 9, 1, 1), g_b, g_gamma, g_beta, None, None, None


class InputPath:
    def __init__(self, cfg=None, device=None, data_mean=None, pfn_params=None, anchors=None,
                 training=True, fused=False, n_lanes=2):
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
        self.mean_prepared = None      # pp_mean_prepare's output, made on first use of the fused path
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
        self.targets_as_list = False   # K3 emits the positives list instead of the dense [B,A,9] tensors
        # fused=True: the step goes through pp_input_path (x never materialised); False: the
        # signature-preserving pp_pillarize -> x -> pp_pfn_scatter sequence
        self.fused = fused
        # the streaming calls of the fused path go through ONE C call per batch (pp_step); False: the Python step
        self.c_step = True
        # host-facing pipeline: copy stream, two device staging buffers, pinned result slots
        self._copy = torch.cuda.Stream(device=self.device)
        # two lanes (main + side stream each) for the *_async calls: consecutive steps alternate lanes, so
        # the latency-bound pillarize stage of step k+1 overlaps the encode stage of step k; the encode
        # stages themselves are ordered (they update the BatchNorm running statistics)
        self._lanes = [(torch.cuda.Stream(device=self.device), torch.cuda.Stream(device=self.device))
                       for _ in range(max(2, int(n_lanes)))]
        self._lane_no = 0
        self._last_encode = None
        self._stage = [None] * len(self._lanes)
        self._slot_free = [None] * len(self._lanes)
        self._result_pin = [None] * 8      # ring of pinned result buffers: a handle's counters stay valid for 8 more steps
        self._step_no = 0

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
        grid = self._grid_struct()
        nbytes = L.pp_pillarize_workspace_bytes(B, T, grid, P)
        ws = _runtime.workspace(nbytes, dev, "pillarize")
        status = _runtime.status_word(dev)
        with _runtime.on_device(dev):
            rc = L.pp_pillarize(points.data_ptr() if T > 0 else None, dt, points.stride(0),
                                points.stride(1), _lib.i64_array(offsets), B, grid, N, P,
                                self.data_mean.data_ptr() if self.data_mean is not None else None,
                                x.data_ptr(), inds.data_ptr(), npil.data_ptr(), status.data_ptr(),
                                ws.data_ptr(), ws.numel(), _runtime.stream_ptr(dev))
        _lib.check(rc, "pp_pillarize")
        return x, inds, npil

    # -- K2 -----------------------------------------------------------------------------------
    def encode(self, x, inds, out=None, return_features=False):
        with torch.no_grad():                       # the streaming path is not differentiated (train through pp_b200.model)
            return self.net(x, inds, return_features=return_features, out=out)

    # -- K1 + K2 fused, x never materialised (pp_input_path) ------------------------------------
    def _grid_struct(self):
        g = self.__dict__.get("_grid_cache")
        if g is None:
            g = self.__dict__["_grid_cache"] = self.cfg.grid()
        return g

    def fused_supported(self, n_sweeps):
        c = self.cfg
        return (1 <= n_sweeps <= _lib.PP_MAX_SWEEPS and c.feature_net_out == 64 and 16 <= c.max_points_per_pillar <= 255
                and c.max_points_per_pillar % 8 == 0 and c.max_pillars % 2 == 0)

    def prepare_mean(self):
        """pp_mean_prepare once per data_mean: the fp16-split tensor-core operand + input moments of the fused
        path's padding pass (pillar_means.pkl is a constant of the dataset)."""
        if self.data_mean is None or self.mean_prepared is not None:
            return self.mean_prepared
        L = _lib.load()
        c = self.cfg
        n = L.pp_mean_prepared_bytes(c.max_pillars, c.max_points_per_pillar)
        buf = torch.empty(n + 256, dtype=torch.uint8, device=self.device)
        buf = buf[(-buf.data_ptr()) % 256:][:n]
        with _runtime.on_device(self.device):
            rc = L.pp_mean_prepare(self.data_mean.data_ptr(), c.max_pillars, c.max_points_per_pillar, buf.data_ptr(), n,
                                   _runtime.stream_ptr(self.device))
        _lib.check(rc, "pp_mean_prepare")
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self._mean_ready = ev
        self.mean_prepared = buf
        return buf

    def pillarize_encode(self, points, offsets, out=None, want_x=False, stages=3):
        """points/offsets as in ``pillarize``.  Returns (canvas, inds, n_pillars[, x]): the canvas of
        ``encode(pillarize(...))`` without writing and re-reading the dense [B,9,P,N] tensor (it is
        only produced when ``want_x``).  ``stages`` = 1 runs only the pillarize stage, 2 only the encode
        stage on the state the pillarize stage left in the workspace of the current stream (same
        ``points`` / ``offsets`` / ``out`` for both calls)."""
        L = _lib.load()
        c = self.cfg
        _runtime.require_cuda(points, "points")
        if points.dtype == torch.float32:
            dt = _lib.PP_F32
        elif points.dtype == torch.float64:
            dt = _lib.PP_F64
        else:
            raise _lib.PPError("points must be float32 or float64")
        B = len(offsets) - 1
        if not self.fused_supported(B):
            raise _lib.PPError("fused input path: 1..%d sweeps, C=64, N<=255 and N%%8==0, P even" % _lib.PP_MAX_SWEEPS)
        T = int(offsets[-1])
        P, N, C = c.max_pillars, c.max_points_per_pillar, c.feature_net_out
        H, W = c.canvas_height, c.canvas_width
        dev = self.device
        o = out or {}
        canvas = o.get("canvas")
        if canvas is None:
            canvas = torch.empty((B, C, H, W), dtype=torch.float32, device=dev)
        if o.get("pillars") is not None:
            x, inds, npil = o["pillars"]
        else:
            x = torch.empty((B, 9, P, N), dtype=torch.float32, device=dev) if want_x else None
            inds = torch.empty((B, P, 3), dtype=torch.int64, device=dev)
            npil = torch.empty(B, dtype=torch.int32, device=dev)
        if not want_x:
            x = None
        net = self.net
        grid = self._grid_struct()
        prep = self.mean_prepared
        if prep is None and self.data_mean is not None:
            prep = self.prepare_mean()
        if prep is not None:
            torch.cuda.current_stream(dev).wait_event(self._mean_ready)
        nbytes = L.pp_input_path_workspace_bytes(B, T, grid, N, P, C, H, W, 0)
        ws = _runtime.workspace(nbytes, dev, "input_path")
        status = _runtime.status_word(dev)
        momentum, eps = _bn_args(net.bn1)
        w = net.conv1.weight
        if not w.is_contiguous():
            w = w.detach().contiguous()
        nbt = net.bn1.num_batches_tracked
        with _runtime.on_device(dev):
            rc = L.pp_input_path(
                points.data_ptr() if T > 0 else None, dt, points.stride(0), points.stride(1),
                _lib.i64_array(offsets), B, grid, N, P,
                self.data_mean.data_ptr() if self.data_mean is not None else None,
                prep.data_ptr() if prep is not None else None, C,
                w.data_ptr(), net.conv1.bias.data_ptr(), net.bn1.weight.data_ptr(),
                net.bn1.bias.data_ptr(), net.bn1.running_mean.data_ptr(),
                net.bn1.running_var.data_ptr(), nbt.data_ptr() if nbt is not None else None,
                1 if net.training else 0, momentum, eps, H, W, canvas.data_ptr(),
                x.data_ptr() if x is not None else None, inds.data_ptr(), npil.data_ptr(),
                status.data_ptr(), ws.data_ptr(), ws.numel(), int(stages), _runtime.stream_ptr(dev))
        _lib.check(rc, "pp_input_path")
        return (canvas, inds, npil, x) if want_x else (canvas, inds, npil)

    def pillarize_encode_train(self, points, offsets):
        """``pillarize_encode`` as a differentiable step: returns (canvas, inds, n_pillars) where ``canvas`` carries
        the autograd graph to ``self.net``'s conv1 / bn1 parameters (pp_input_path_backward: the gradients of the
        reference's PPFeatureNet + PPScatter from the compact per-point state; the dense x is never built in either
        direction).  The forward's workspace is kept alive by the graph until backward has run."""
        net = self.net
        return _InputPathFunction.apply(net.conv1.weight, net.conv1.bias, net.bn1.weight, net.bn1.bias, self, points,
                                        list(offsets))

    # -- K3 -----------------------------------------------------------------------------------
    def targets(self, gt_dev, gt_offsets, out=None):
        """gt_dev: dict of CUDA tensors corners [Gt,4,2], centers [Gt,3] (image space), wlh [Gt,3],
        yaw [Gt] (float64), cls [Gt] (int32).  With ``self.targets_as_list`` the first return value is a
        ``box_utils.Positives`` list (for ``loss.PPLoss``) and the second is None: no dense tensor is written."""
        a = self.ensure_anchors()
        return assign_targets(a, gt_dev["corners"], gt_dev["centers"], gt_dev["wlh"], gt_dev["yaw"],
                              gt_dev["cls"], gt_offsets, self.cfg.num_classes,
                              self.cfg.iou_pos_thresh, out=out, as_list=self.targets_as_list)

    # -- host-facing step -----------------------------------------------------------------------
    @staticmethod
    def _blob_layout(T, ncol, Gt, F=0):
        """Byte offsets of the sections of one packed batch (each 256-byte aligned).  F > 0 appends the
        per-file transforms and file offsets of a batch that is aggregated on the device."""
        al = lambda v: (v + 255) // 256 * 256
        sizes = [("points", max(T, 1) * ncol * 4), ("corners", max(Gt, 1) * 8 * 8), ("centers", max(Gt, 1) * 3 * 8),
                 ("wlh", max(Gt, 1) * 3 * 8), ("yaw", max(Gt, 1) * 8), ("cls", max(Gt, 1) * 4)]
        if F > 0:
            sizes += [("xforms", F * 12 * 8), ("file_offsets", (F + 1) * 8)]
        off, o = {}, 0
        for name, n in sizes:
            off[name] = (o, n)
            o = al(o + n)
        return off, o

    @staticmethod
    def _blob_views(blob, off, T, ncol, Gt):
        v = lambda name, dt: blob[off[name][0]:off[name][0] + off[name][1]].view(dt)
        views = {"points": v("points", torch.float32).view(max(T, 1), ncol),
                 "corners": v("corners", torch.float64).view(max(Gt, 1), 8),
                 "centers": v("centers", torch.float64).view(max(Gt, 1), 3),
                 "wlh": v("wlh", torch.float64).view(max(Gt, 1), 3),
                 "yaw": v("yaw", torch.float64), "cls": v("cls", torch.int32)}
        if "xforms" in off:
            views["xforms"] = v("xforms", torch.float64).view(-1, 12)
            views["file_offsets"] = v("file_offsets", torch.int64)
        return views

    # -- sweep aggregation in front of K1 (pp_aggregate_sweeps; data/dataset.py:54-88) -------------------
    def aggregate(self, d_points, file_offsets, transforms, min_dist=0.001, want_kept=False):
        """In place on device rows ``d_points [T, S>=3]`` float32: every file's points go through its 4x4
        (or 3x4) float64 ``transmat`` (dataset.py:78) and the SDK's ``remove_close(min_dist)``; dropped points
        get an out-of-range sentinel that the pillarize stage filters.  ``file_offsets`` (F+1 row offsets) and
        ``transforms`` may be host sequences or device tensors.  Returns the per-file kept counts if asked."""
        L = _lib.load()
        _runtime.require_cuda(d_points, "d_points")
        if d_points.dtype != torch.float32 or d_points.dim() != 2 or not d_points.is_contiguous():
            raise _lib.PPError("aggregate: d_points must be a contiguous float32 [T, S] tensor")
        dev = d_points.device
        fo = torch.as_tensor(file_offsets, dtype=torch.int64).to(dev).contiguous()
        if isinstance(transforms, torch.Tensor) and transforms.is_cuda:
            xf = transforms.to(torch.float64).reshape(fo.numel() - 1, -1, 4)[:, :3, :].contiguous()
        else:
            xf = torch.from_numpy(np.ascontiguousarray(
                np.asarray(transforms, dtype=np.float64).reshape(fo.numel() - 1, -1, 4)[:, :3, :])).to(dev)
        F = fo.numel() - 1
        kept = torch.zeros(F, dtype=torch.int32, device=dev) if want_kept else None
        with _runtime.on_device(dev):
            rc = L.pp_aggregate_sweeps(d_points.data_ptr(), d_points.shape[0], d_points.shape[1], fo.data_ptr(), F,
                                       xf.data_ptr(), float(min_dist), kept.data_ptr() if kept is not None else None,
                                       _runtime.stream_ptr(dev))
        _lib.check(rc, "pp_aggregate_sweeps")
        return kept

    def pack_host_batch(self, sweeps, gts, transforms=None, min_dist=0.001):
        """Pack a list of float32 [n_i, >=4] sweeps and a list of GT dicts (centers/wlh/yaw/cls in
        canvas space, as the reference's box pickles hold them) into ONE pinned host buffer
        (points | GT corners | centres | wlh | yaw | class ids), so that a step needs a single
        host-to-device copy and no device-side repacking.

        With ``transforms``: ``sweeps[i]`` is the LIST of raw lidar files ([n,5] float32, in the order
        dataset.py:58-85 visits them) that sample i aggregates and ``transforms[i]`` the matching list of
        4x4 float64 ``transmat``; the transform and ``remove_close(min_dist)`` then run on the device right
        after the upload (``aggregate``), nothing is transformed or compacted on the host."""
        files, xf, foffs = sweeps, None, None
        if transforms is not None:
            files, xf, foffs, offs = [], [], [0], [0]
            for group, mats in zip(sweeps, transforms):
                group = group if isinstance(group, (list, tuple)) else [group]
                mats = mats if isinstance(mats, (list, tuple)) else [mats]
                if len(group) != len(mats):
                    raise _lib.PPError("pack_host_batch: one transform per lidar file")
                for f, m in zip(group, mats):
                    files.append(f)
                    xf.append(np.asarray(m, dtype=np.float64).reshape(-1, 4)[:3, :])
                    foffs.append(foffs[-1] + int(f.shape[0]))
                offs.append(foffs[-1])
        else:
            offs = [0]
            for s in sweeps:
                offs.append(offs[-1] + int(s.shape[0]))
        ncol = int(files[0].shape[1])
        goffs = [0]
        for g in gts:
            goffs.append(goffs[-1] + int(len(g["yaw"])))
        T, Gt = offs[-1], goffs[-1]
        F = len(xf) if xf is not None else 0
        layout, nbytes = self._blob_layout(T, ncol, Gt, F)
        blob = torch.zeros(nbytes, dtype=torch.uint8).pin_memory()
        hv = self._blob_views(blob, layout, T, ncol, Gt)
        lo = 0
        for f in files:
            hv["points"][lo:lo + f.shape[0]] = torch.as_tensor(f, dtype=torch.float32)
            lo += int(f.shape[0])
        if F:
            hv["xforms"][:] = torch.from_numpy(np.stack(xf).reshape(F, 12))
            hv["file_offsets"][:] = torch.tensor(foffs, dtype=torch.int64)
        for g, lo, hi in zip(gts, goffs[:-1], goffs[1:]):
            if hi == lo:
                continue
            cen, cor = gt_to_image_space(g, self.cfg.canvas_height)
            hv["corners"][lo:hi] = torch.from_numpy(cor.reshape(-1, 8))
            hv["centers"][lo:hi] = torch.from_numpy(cen)
            hv["wlh"][lo:hi] = torch.from_numpy(np.asarray(g["wlh"], dtype=np.float64))
            hv["yaw"][lo:hi] = torch.from_numpy(np.asarray(g["yaw"], dtype=np.float64))
            hv["cls"][lo:hi] = torch.from_numpy(np.asarray(g["cls"], dtype=np.int32))
        return {"blob": blob, "layout": layout, "ncol": ncol, "offsets": offs, "gt_offsets": goffs,
                "points": hv["points"], "host": hv, "n_files": F, "min_dist": float(min_dist)}

    def upload(self, batch, slot=None):
        """One host-to-device copy of a packed batch; returns (d_points, gt_dev) views of the device
        copy.  ``slot`` selects one of the two resident staging buffers (step_host_async alternates)."""
        dev = self.device
        T, Gt = batch["offsets"][-1], batch["gt_offsets"][-1]
        n = batch["blob"].numel()
        if slot is None:
            dblob = torch.empty(n, dtype=torch.uint8, device=dev)
        else:
            dblob = self._stage[slot]
            if dblob is None or dblob.numel() < n:
                dblob = self._stage[slot] = torch.empty(n, dtype=torch.uint8, device=dev)
        dblob[:n].copy_(batch["blob"], non_blocking=True)
        dv = self._blob_views(dblob, batch["layout"], T, batch["ncol"], Gt)
        gt_dev = {k: dv[k] for k in ("corners", "centers", "wlh", "yaw", "cls")}
        d_pts = dv["points"][:max(T, 1)]
        if batch.get("n_files"):                     # aggregate on the stream the copy was issued on
            with _runtime.on_device(dev):
                rc = _lib.load().pp_aggregate_sweeps(d_pts.data_ptr(), T, batch["ncol"], dv["file_offsets"].data_ptr(),
                                                     batch["n_files"], dv["xforms"].data_ptr(), batch["min_dist"], None,
                                                     _runtime.stream_ptr(dev))
            _lib.check(rc, "pp_aggregate_sweeps")
        return d_pts, gt_dev

    def _run(self, d_pts, offsets, gt_dev, gt_offsets, o, side=None, ordered=False):
        main = torch.cuda.current_stream(self.device)
        side = self._side if side is None else side
        if self.overlap_targets:
            self.ensure_anchors()
            side.wait_stream(main)
            with torch.cuda.stream(side):
                cls, reg, top, counts = self.targets(gt_dev, gt_offsets, out=o.get("targets"))
            canvas, npil = self._k12(d_pts, offsets, o, main, ordered)
            main.wait_stream(side)
            held = (cls.anchor, cls.cls, cls.reg, cls.offsets, top, counts) if reg is None else (cls, reg, top, counts)
            for t in held:
                t.record_stream(main)
        else:
            canvas, npil = self._k12(d_pts, offsets, o, main, ordered)
            cls, reg, top, counts = self.targets(gt_dev, gt_offsets, out=o.get("targets"))
        return canvas, cls, reg, npil, counts

    def _k12(self, d_pts, offsets, o, main, ordered):
        """Pillarize + encode on the current stream.  ``ordered``: called from a lane; the encode stage
        waits for the encode stage of the previous step (running statistics are read-modify-write)."""
        fused = self.fused and self.fused_supported(len(offsets) - 1)
        if fused:
            if o.get("pillars") is None or o.get("canvas") is None:
                o = dict(o)
                c = self.cfg
                B = len(offsets) - 1
                if o.get("canvas") is None:
                    o["canvas"] = torch.empty((B, c.feature_net_out, c.canvas_height, c.canvas_width),
                                              dtype=torch.float32, device=self.device)
                if o.get("pillars") is None:
                    o["pillars"] = (None, torch.empty((B, c.max_pillars, 3), dtype=torch.int64, device=self.device),
                                    torch.empty(B, dtype=torch.int32, device=self.device))
            if ordered:
                self.pillarize_encode(d_pts, offsets, out=o, stages=1)
                if self._last_encode is not None:
                    main.wait_event(self._last_encode)
                canvas, inds, npil = self.pillarize_encode(d_pts, offsets, out=o, stages=2)
            else:
                canvas, inds, npil = self.pillarize_encode(d_pts, offsets, out=o)
        else:
            x, inds, npil = self.pillarize(d_pts, offsets, out=o.get("pillars"))
            if ordered and self._last_encode is not None:
                main.wait_event(self._last_encode)
            canvas = self.encode(x, inds, out=o.get("canvas"))
        if ordered:
            self._last_encode = torch.cuda.Event()
            self._last_encode.record(main)
        return canvas, npil

    # -- one C call per batch (pp_step) ----------------------------------------------------------------------
    def _lane_ctx(self, k):
        """Per-lane state of the C step driver: streams, pre-created events (raw handles), workspaces."""
        ctx = self.__dict__.setdefault("_cstep_lanes", {}).get(k)
        if ctx is None:
            main, side = self._lanes[k]

            def event():
                e = torch.cuda.Event()
                e.record(main)                       # materialises the cudaEvent_t; pp_step re-records it
                return e
            ctx = {"main": main, "side": side, "fork": event(), "join": event(), "ws_in": None, "ws_tg": None}
            self._cstep_lanes[k] = ctx
        return ctx

    def _ring_ctx(self, i):
        """Per in-flight-step state: completion / upload / encode-order events and the pinned counters."""
        ring = self.__dict__.setdefault("_cstep_ring", {})
        ctx = ring.get(i)
        if ctx is None:
            def event():
                e = torch.cuda.Event()
                e.record(self._copy)
                return e
            ctx = {"done": event(), "ready": event(), "encode": event(), "pin": None, "plan": _lib.PPStepPlan()}
            ring[i] = ctx
        return ctx

    def _step_c(self, batch, d_pts, offsets, gt_dev, gt_offsets, out):
        """Issue one batch through pp_step.  ``batch`` (a packed host batch) or ``d_pts`` / ``gt_dev`` (device)."""
        L = _lib.load()
        c = self.cfg
        dev = self.device
        B = len(offsets) - 1
        P, N, C = c.max_pillars, c.max_points_per_pillar, c.feature_net_out
        H, W = c.canvas_height, c.canvas_width
        net = self.net
        anchors = self.ensure_anchors()
        A = anchors.A
        k = self._lane_no % len(self._lanes)
        self._lane_no += 1
        lane = self._lane_ctx(k)
        no = self._step_no
        self._step_no += 1
        R = 8
        ring = self._ring_ctx(no % R)
        slot = no % len(self._stage)
        prep = self.mean_prepared
        if prep is None and self.data_mean is not None:
            prep = self.prepare_mean()
            torch.cuda.current_stream(dev).synchronize()
        T = int(offsets[-1])
        Gt = int(gt_offsets[-1])
        o = out or {}
        canvas = o.get("canvas")
        if canvas is None:
            canvas = torch.empty((B, C, H, W), dtype=torch.float32, device=dev)
        if o.get("pillars") is not None:
            _, inds, npil = o["pillars"]
        else:
            inds = torch.empty((B, P, 3), dtype=torch.int64, device=dev)
            npil = torch.empty(B, dtype=torch.int32, device=dev)
        pos = None
        if self.targets_as_list:
            from .box_utils import Positives
            cap = 4096 * B
            pos = Positives(torch.empty(cap, dtype=torch.int32, device=dev), torch.empty((cap, 9), dtype=torch.float32, device=dev),
                            torch.empty((cap, 9), dtype=torch.float32, device=dev),
                            torch.empty(B + 1, dtype=torch.int32, device=dev), B, A)
            cls = reg = None
        elif o.get("targets") is not None:
            cls, reg = o["targets"]
        else:
            cls = torch.empty((B, A, c.num_classes), dtype=torch.float32, device=dev)
            reg = torch.empty((B, A, 9), dtype=torch.float32, device=dev)
        top = torch.empty(max(Gt, 1), dtype=torch.int32, device=dev)
        counts = torch.empty((B, 4), dtype=torch.int32, device=dev)
        grid = self._grid_struct()
        n_in = L.pp_input_path_workspace_bytes(B, T, grid, N, P, C, H, W, 0)
        if lane["ws_in"] is None or lane["ws_in"].numel() < n_in:
            lane["ws_in"] = torch.empty(n_in, dtype=torch.uint8, device=dev)
        n_tg = L.pp_assign_targets_workspace_bytes(B, A, Gt, None)
        if lane["ws_tg"] is None or lane["ws_tg"].numel() < n_tg:
            lane["ws_tg"] = torch.empty(n_tg, dtype=torch.uint8, device=dev)
        pin = ring["pin"]
        if pin is None or pin.numel() < 5 * B + 1:
            pin = ring["pin"] = torch.empty(5 * B + 1, dtype=torch.int32).pin_memory()
        momentum, eps = _bn_args(net.bn1)
        w = net.conv1.weight
        if not w.is_contiguous():
            w = w.detach().contiguous()
        nbt = net.bn1.num_batches_tracked
        status = _runtime.status_word(dev)
        pl = ring["plan"]
        ptr = lambda t: t.data_ptr() if (t is not None and t.numel() > 0) else None
        pl.stream_main, pl.stream_side, pl.stream_copy = lane["main"].cuda_stream, lane["side"].cuda_stream, self._copy.cuda_stream
        pl.ev_fork, pl.ev_join = lane["fork"].cuda_event, lane["join"].cuda_event
        pl.ev_done, pl.ev_encode_done = ring["done"].cuda_event, ring["encode"].cuda_event
        pl.ev_prev_encode = self._last_encode.cuda_event if self._last_encode is not None else None
        if batch is not None:
            n = batch["blob"].numel()
            dblob = self._stage[slot]
            if dblob is None or dblob.numel() < n:
                dblob = self._stage[slot] = torch.empty(n, dtype=torch.uint8, device=dev)
            dv = self._blob_views(dblob, batch["layout"], T, batch["ncol"], Gt)
            gt_dev = {kk: dv[kk] for kk in ("corners", "centers", "wlh", "yaw", "cls")}
            d_pts = dv["points"][:max(T, 1)]
            pl.h_blob, pl.d_blob, pl.blob_bytes = batch["blob"].data_ptr(), dblob.data_ptr(), n
            pl.ev_ready = ring["ready"].cuda_event
            if self._slot_free[slot] is None:
                e = torch.cuda.Event()
                e.record(lane["main"])
                self._slot_free[slot] = e
            pl.ev_slot_free = self._slot_free[slot].cuda_event
            pl.n_files = int(batch.get("n_files") or 0)
            if pl.n_files:
                pl.d_file_offsets, pl.d_file_xforms = dv["file_offsets"].data_ptr(), dv["xforms"].data_ptr()
                pl.min_dist = float(batch["min_dist"])
            # the upload is ordered after whatever the caller queued on its current stream
            self._copy.wait_stream(torch.cuda.current_stream(dev))
        else:
            pl.h_blob = pl.d_blob = pl.ev_ready = pl.ev_slot_free = None
            pl.blob_bytes, pl.n_files = 0, 0
            lane["main"].wait_stream(torch.cuda.current_stream(dev))
        pl.d_points = d_pts.data_ptr() if T > 0 else None
        pl.total_points, pl.point_cols = T, int(d_pts.stride(0))
        pl.h_sweep_offsets = ctypes.cast(_lib.i64_array(offsets), ctypes.c_void_p)
        pl.n_sweeps = B
        pl.grid = grid
        pl.max_points_per_pillar, pl.max_pillars = N, P
        pl.d_data_mean = self.data_mean.data_ptr() if self.data_mean is not None else None
        pl.d_mean_prepared = prep.data_ptr() if prep is not None else None
        pl.C = C
        pl.d_conv_w, pl.d_conv_b = w.data_ptr(), net.conv1.bias.data_ptr()
        pl.d_bn_w, pl.d_bn_b = net.bn1.weight.data_ptr(), net.bn1.bias.data_ptr()
        pl.d_running_mean, pl.d_running_var = net.bn1.running_mean.data_ptr(), net.bn1.running_var.data_ptr()
        pl.d_num_batches_tracked = nbt.data_ptr() if nbt is not None else None
        pl.training, pl.momentum, pl.eps = (1 if net.training else 0), momentum, eps
        pl.canvas_h, pl.canvas_w = H, W
        pl.d_canvas, pl.d_indices, pl.d_num_pillars = canvas.data_ptr(), inds.data_ptr(), npil.data_ptr()
        pl.d_ws_input, pl.ws_input_bytes = lane["ws_in"].data_ptr(), lane["ws_in"].numel()
        pl.d_a_corners, pl.d_a_centers = anchors.corners.data_ptr(), anchors.centers.data_ptr()
        pl.d_a_wlh, pl.d_a_yaw, pl.d_anchor_index, pl.A = anchors.wlh.data_ptr(), anchors.yaw.data_ptr(), anchors.index.data_ptr(), A
        pl.d_g_corners, pl.d_g_centers = ptr(gt_dev["corners"]) if Gt else None, ptr(gt_dev["centers"]) if Gt else None
        pl.d_g_wlh, pl.d_g_yaw, pl.d_g_cls = (ptr(gt_dev["wlh"]), ptr(gt_dev["yaw"]), ptr(gt_dev["cls"])) if Gt else (None, None, None)
        pl.h_gt_offsets = ctypes.cast(_lib.i64_array(gt_offsets), ctypes.c_void_p)
        pl.num_classes, pl.pos_thresh = int(c.num_classes), float(c.iou_pos_thresh)
        pl.d_cls, pl.d_reg = ptr(cls), ptr(reg)
        pl.d_top_anchor, pl.d_counts = top.data_ptr(), counts.data_ptr()
        if pos is not None:
            pl.d_pos_anchor, pl.d_pos_cls, pl.d_pos_reg = pos.anchor.data_ptr(), pos.cls.data_ptr(), pos.reg.data_ptr()
            pl.d_pos_offsets, pl.pos_capacity = pos.offsets.data_ptr(), pos.anchor.numel()
        else:
            pl.d_pos_anchor = pl.d_pos_cls = pl.d_pos_reg = pl.d_pos_offsets = None
            pl.pos_capacity = 0
        pl.d_ws_targets, pl.ws_targets_bytes = lane["ws_tg"].data_ptr(), lane["ws_tg"].numel()
        pl.d_status, pl.h_counters = status.data_ptr(), pin.data_ptr()
        with _runtime.on_device(dev):
            rc = L.pp_step(ctypes.byref(pl))
        _lib.check(rc, "pp_step")
        self._last_encode = ring["encode"]
        res = (canvas, pos if pos is not None else cls, reg, npil, counts)
        h = StepHandle(res, pin, B, ring["done"])
        h._keep = (top, d_pts, gt_dev)
        return h

    def _next_lane(self):
        lane = self._lanes[self._lane_no % len(self._lanes)]
        self._lane_no += 1
        return lane

    def step_host(self, batch, out=None, check=True):
        """One pass of the whole path from a pinned HOST batch (``pack_host_batch``): H2D copy,
        pillarize, PFN + scatter, target assignment.  Outputs stay on the device, where the
        backbone and the loss consume them; returns (canvas, cls, reg, n_pillars, counts).
        ``check``: read the device status word afterwards (synchronises) and raise on a set bit, as the
        reference raises on an out-of-canvas scatter index; the streaming ``step_host_async`` reports it
        through ``StepHandle.counters()`` instead."""
        d_pts, gt_dev = self.upload(batch)
        res = self._run(d_pts, batch["offsets"], gt_dev, batch["gt_offsets"], out or {})
        if check:
            _runtime.check_status(self.device, "step_host")
        return res

    def step_host_async(self, batch, out=None):
        """Pipelined form of ``step_host`` for a streaming loop: the H2D copy goes to a copy stream
        into one of two staging buffers, the kernels to one of two lanes (so the copy and the
        pillarize stage of step k+1 overlap the encode stage of step k), the step's counters come
        back through a pinned buffer, and nothing blocks the host.  Consecutive calls must not share
        ``out`` buffers.  Returns a ``StepHandle``; ``handle.counters()`` waits for this step only and must
        be called before eight further steps have been issued (the pinned result buffers form a ring)."""
        if self.fused and self.c_step and self.fused_supported(len(batch["offsets"]) - 1) and batch["points"].dtype == torch.float32:
            return self._step_c(batch, None, batch["offsets"], None, batch["gt_offsets"], out)
        dev = self.device
        slot = self._step_no % len(self._stage)
        self._step_no += 1
        main, side = self._next_lane()
        if self._slot_free[slot] is not None:
            self._copy.wait_event(self._slot_free[slot])     # kernels that read this staging buffer two steps ago
        self._copy.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(self._copy):
            d_pts, gt_dev = self.upload(batch, slot)
            ready = torch.cuda.Event()
            ready.record(self._copy)
        with torch.cuda.stream(main):
            main.wait_event(ready)
            res = self._run(d_pts, batch["offsets"], gt_dev, batch["gt_offsets"], out or {}, side=side, ordered=True)
            self._slot_free[slot] = torch.cuda.Event()
            self._slot_free[slot].record(main)
            B = len(batch["offsets"]) - 1
            ring = (self._step_no - 1) % len(self._result_pin)
            pin = self._result_pin[ring]
            if pin is None or pin.numel() < 5 * B + 1:
                pin = self._result_pin[ring] = torch.empty(5 * B + 1, dtype=torch.int32).pin_memory()
            pin[:B].copy_(res[3], non_blocking=True)
            pin[B:5 * B].view(B, 4).copy_(res[4], non_blocking=True)
            pin[5 * B:5 * B + 1].copy_(_runtime.status_word(dev), non_blocking=True)
            done = torch.cuda.Event()
            done.record(main)
        return StepHandle(res, pin, B, done)

    def step_device_async(self, d_pts, offsets, gt_dev, gt_offsets, out=None):
        """``step_device`` on one of the two lanes (see ``step_host_async``); inputs must be ready on the
        current stream.  Returns a ``StepHandle`` (``wait()`` orders the current stream after the step)."""
        if (self.fused and self.c_step and self.fused_supported(len(offsets) - 1) and d_pts.dtype == torch.float32
                and d_pts.stride(1) == 1):
            return self._step_c(None, d_pts, offsets, gt_dev, gt_offsets, out)
        dev = self.device
        main, side = self._next_lane()
        main.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(main):
            res = self._run(d_pts, offsets, gt_dev, gt_offsets, out or {}, side=side, ordered=True)
            done = torch.cuda.Event()
            done.record(main)
        return StepHandle(res, None, len(offsets) - 1, done)

    def step_device(self, d_pts, offsets, gt_dev, gt_offsets, out=None, check=False):
        """Same pass with inputs already resident in HBM.  ``check`` as in ``step_host`` (default off: the caller
        polls ``check_status`` itself)."""
        res = self._run(d_pts, offsets, gt_dev, gt_offsets, out or {})
        if check:
            _runtime.check_status(self.device, "step_device")
        return res

    def check_status(self):
        """Synchronising read of the device status word; raises on a set bit and clears it."""
        _runtime.check_status(self.device, "input path")
