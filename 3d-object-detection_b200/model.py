"""Drop-ins for the first two modules of the reference network (model/model.py:13-62).

``PPFeatureNet`` and ``PPScatter`` keep the reference's constructors, sub-module / parameter
names (``conv1.weight [C,9,1,1]``, ``conv1.bias``, ``bn1.*``) and ``forward`` signatures, so
checkpoints written by the reference's train.py load unchanged and ``PPModel`` can use them as
is.  ``forward`` runs the sm_100a kernels of libpp_b200.so; there is no eager fallback: CPU
tensors raise.  ``PPFeatureScatter`` is the fused module (x, inds) -> canvas.

Forward only in this round: the custom autograd function raises in backward.
"""
import torch
import torch.nn as nn

from . import _lib, _runtime
from .config import cfg as _cfg


def _bn_args(bn):
    if bn.running_mean is None or bn.running_var is None:
        raise _lib.PPError("BatchNorm2d without running statistics is not supported")
    momentum = 0.1 if bn.momentum is None else float(bn.momentum)
    if bn.momentum is None and bn.training:
        raise _lib.PPError("momentum=None (cumulative average) is not supported")
    return momentum, float(bn.eps)


def _prep(x, conv, bn):
    _runtime.require_cuda(x, "x")
    if x.dtype != torch.float32:
        raise _lib.PPError("x must be float32 (the reference feeds .float() tensors)")
    if conv.kernel_size != (1, 1) or conv.bias is None:
        raise _lib.PPError("conv1 must be a 1x1 convolution with bias")
    return x.contiguous()


class _PFNForward(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, conv_w, conv_b, bn_w, bn_b, module):
        out = module._run(x)
        return out

    @staticmethod
    def backward(ctx, *grads):  # pragma: no cover
        raise NotImplementedError(
            "PPFeatureNet backward is not built yet (SURVEY.md 8(f) row N1); run under "
            "torch.no_grad() or detach the input")


class PPFeatureNet(nn.Module):
    """model/model.py:13-40: [B,D,P,N] -> conv1 1x1 -> ReLU -> bn1 -> max over N -> [B,C,P]."""

    def __init__(self, in_channels, out_channels):
        super(PPFeatureNet, self).__init__()
        self.conv1 = nn.Conv2d(in_channels, out_channels, kernel_size=1)
        self.bn1 = nn.BatchNorm2d(out_channels)

    def _run(self, x):
        L = _lib.load()
        x = _prep(x, self.conv1, self.bn1)
        B, D, P, N = x.shape
        C = self.conv1.out_channels
        dev = x.device
        momentum, eps = _bn_args(self.bn1)
        out = torch.empty((B, C, P), dtype=torch.float32, device=dev)
        nbytes = L.pp_pfn_workspace_bytes(B, P, C, 0, 0)
        ws = _runtime.workspace(nbytes, dev, "pfn")
        w = self.conv1.weight.detach().reshape(C, D).contiguous()
        nbt = self.bn1.num_batches_tracked
        with torch.cuda.device(dev):
            rc = L.pp_pfn_forward(
                x.data_ptr(), B, D, P, N, C, w.data_ptr(), self.conv1.bias.detach().data_ptr(),
                self.bn1.weight.detach().data_ptr(), self.bn1.bias.detach().data_ptr(),
                self.bn1.running_mean.data_ptr(), self.bn1.running_var.data_ptr(),
                nbt.data_ptr() if nbt is not None else None, 1 if self.training else 0, momentum,
                eps, out.data_ptr(), ws.data_ptr(), ws.numel(), _runtime.stream_ptr(dev))
        _lib.check(rc, "pp_pfn_forward")
        return out

    def forward(self, x):
        if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters())):
            return _PFNForward.apply(x, self.conv1.weight, self.conv1.bias, self.bn1.weight,
                                     self.bn1.bias, self)
        return self._run(x)


class PPScatter(nn.Module):
    """model/model.py:42-62: [B,C,P] + inds [B,P,3] (int64: flag, x, y) -> canvas [B,C,H,W]."""

    def __init__(self, device=None, canvas_height=None, canvas_width=None):
        super(PPScatter, self).__init__()
        self.device = device
        self.canvas_height = int(_cfg.canvas_height if canvas_height is None else canvas_height)
        self.canvas_width = int(_cfg.canvas_width if canvas_width is None else canvas_width)

    def forward(self, x, inds):
        L = _lib.load()
        _runtime.require_cuda(x, "x")
        _runtime.require_cuda(inds, "inds")
        if x.dtype != torch.float32 or inds.dtype != torch.int64:
            raise _lib.PPError("x must be float32 and inds int64")
        x = x.contiguous()
        inds = inds.contiguous()
        B, C, P = x.shape
        H, W = self.canvas_height, self.canvas_width
        dev = x.device
        out = torch.empty((B, C, H, W), dtype=torch.float32, device=dev)
        ws = _runtime.workspace(B * H * W * 4 + 1024, dev, "scatter")
        status = _runtime.status_word(dev)
        with torch.cuda.device(dev):
            rc = L.pp_scatter(x.data_ptr(), inds.data_ptr(), B, C, P, H, W, out.data_ptr(),
                              status.data_ptr(), ws.data_ptr(), ws.numel(), _runtime.stream_ptr(dev))
        _lib.check(rc, "pp_scatter")
        return out


class PPFeatureScatter(nn.Module):
    """PPFeatureNet + PPScatter in one pass over x: (x [B,D,P,N], inds [B,P,3]) -> [B,C,H,W].
    Holds the same ``conv1`` / ``bn1`` sub-modules, so a PPFeatureNet state_dict loads."""

    def __init__(self, in_channels, out_channels, canvas_height=None, canvas_width=None):
        super(PPFeatureScatter, self).__init__()
        self.conv1 = nn.Conv2d(in_channels, out_channels, kernel_size=1)
        self.bn1 = nn.BatchNorm2d(out_channels)
        self.canvas_height = int(_cfg.canvas_height if canvas_height is None else canvas_height)
        self.canvas_width = int(_cfg.canvas_width if canvas_width is None else canvas_width)

    @torch.no_grad()
    def forward(self, x, inds, return_features=False, out=None):
        L = _lib.load()
        x = _prep(x, self.conv1, self.bn1)
        _runtime.require_cuda(inds, "inds")
        inds = inds.contiguous()
        B, D, P, N = x.shape
        C = self.conv1.out_channels
        H, W = self.canvas_height, self.canvas_width
        dev = x.device
        momentum, eps = _bn_args(self.bn1)
        canvas = out if out is not None else torch.empty((B, C, H, W), dtype=torch.float32, device=dev)
        feat = torch.empty((B, C, P), dtype=torch.float32, device=dev) if return_features else None
        nbytes = L.pp_pfn_workspace_bytes(B, P, C, H, W)
        ws = _runtime.workspace(nbytes, dev, "pfn")
        status = _runtime.status_word(dev)
        w = self.conv1.weight.detach().reshape(C, D).contiguous()
        nbt = self.bn1.num_batches_tracked
        with torch.cuda.device(dev):
            rc = L.pp_pfn_scatter(
                x.data_ptr(), inds.data_ptr(), B, D, P, N, C, w.data_ptr(),
                self.conv1.bias.detach().data_ptr(), self.bn1.weight.detach().data_ptr(),
                self.bn1.bias.detach().data_ptr(), self.bn1.running_mean.data_ptr(),
                self.bn1.running_var.data_ptr(), nbt.data_ptr() if nbt is not None else None,
                1 if self.training else 0, momentum, eps, H, W, canvas.data_ptr(),
                feat.data_ptr() if feat is not None else None, status.data_ptr(), ws.data_ptr(),
                ws.numel(), _runtime.stream_ptr(dev))
        _lib.check(rc, "pp_pfn_scatter")
        return (canvas, feat) if return_features else canvas
