"""Drop-ins for the first two modules of the reference network (model/model.py:13-62).

``PPFeatureNet`` and ``PPScatter`` keep the reference's constructors, sub-module / parameter
names (``conv1.weight [C,9,1,1]``, ``conv1.bias``, ``bn1.*``) and ``forward`` signatures, so
checkpoints written by the reference's train.py load unchanged and ``PPModel`` can use them as
is.  ``forward`` runs the sm_100a kernels of libpp_b200.so; there is no eager fallback: CPU
tensors raise.  ``PPFeatureScatter`` is the fused module (x, inds) -> canvas.

Training: the three modules are differentiable with respect to conv1.weight/bias and bn1.weight/bias
(pp_pfn_backward, csrc/pfn_bwd.cu -- SURVEY.md 8(f) N1), which is what train.py:147 needs; the gradient with
respect to the input tensor x is not built (x is data in train.py) and asking for it raises.

Checkpoints: parameter names and shapes are the reference's (``conv1.*``, ``bn1.*``; ``strict=True`` loads), but
the per-slot normalisation constant ``pillar_means.pkl`` depends on the pillarizer's pillar order and must be
regenerated with ``pp_b200.make_means``; a checkpoint trained against the reference's means
needs re-validation or fine-tuning.
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


def _pfn_backward(x, conv_w, conv_b, bn_w, running_mean, running_var, training, eps, grad, inds, H, W):
    """pp_pfn_backward: (grad conv1.weight [C,D,1,1], grad conv1.bias, grad bn1.weight, grad bn1.bias).
    ``grad`` is dL/d(features) [B,C,P] (inds None) or the canvas gradient [B,C,H,W] with inds [B,P,3]."""
    L = _lib.load()
    B, D, P, N = x.shape
    C = conv_w.shape[0]
    dev = x.device
    grad = grad.contiguous()
    if grad.dtype != torch.float32:
        raise _lib.PPError("PPFeatureNet backward: the incoming gradient must be float32")
    g_w = torch.empty((C, D), dtype=torch.float32, device=dev)
    g_b = torch.empty(C, dtype=torch.float32, device=dev)
    g_gamma = torch.empty(C, dtype=torch.float32, device=dev)
    g_beta = torch.empty(C, dtype=torch.float32, device=dev)
    nbytes = L.pp_pfn_backward_workspace_bytes(B, P, C)
    if nbytes == 0:
        raise _lib.PPError("PPFeatureNet backward supports in_channels 9, out_channels 64")
    ws = _runtime.workspace(nbytes, dev, "pfn_bwd")
    w2 = conv_w.reshape(C, D).contiguous()
    with torch.cuda.device(dev):
        rc = L.pp_pfn_backward(
            x.data_ptr(), B, D, P, N, C, w2.data_ptr(), conv_b.data_ptr(), bn_w.data_ptr(),
            running_mean.data_ptr() if running_mean is not None else None,
            running_var.data_ptr() if running_var is not None else None, 1 if training else 0, eps,
            grad.data_ptr(), inds.data_ptr() if inds is not None else None, H, W, g_w.data_ptr(), g_b.data_ptr(),
            g_gamma.data_ptr(), g_beta.data_ptr(), ws.data_ptr(), ws.numel(), _runtime.stream_ptr(dev))
    _lib.check(rc, "pp_pfn_backward")
    return g_w.view(C, D, 1, 1), g_b, g_gamma, g_beta


def _no_input_grad(x):
    if x.requires_grad:
        raise _lib.PPError("the gradient with respect to the pillar tensor x is not built (it is input data in the "
                           "reference's train.py); detach x")


class _PFNFunction(torch.autograd.Function):
    """forward: module._run(x) (pp_pfn_forward / pp_pfn_scatter); backward: pp_pfn_backward.  The eval-mode
    running statistics are copied at forward time because the backward normalises with them."""

    @staticmethod
    def forward(ctx, x, conv_w, conv_b, bn_w, bn_b, module, inds):
        _no_input_grad(x)
        ctx.training = bool(module.training)
        ctx.eps = float(module.bn1.eps)
        ctx.canvas = inds is not None
        if ctx.training:
            rm = rv = None
        else:
            rm, rv = module.bn1.running_mean.clone(), module.bn1.running_var.clone()
        out = module._run(x, inds) if ctx.canvas else module._run(x)
        ctx.hw = (out.shape[2], out.shape[3]) if ctx.canvas else (0, 0)
        ctx.stats = (rm, rv)
        ctx.save_for_backward(x, conv_w, conv_b, bn_w, inds)
        return out

    @staticmethod
    def backward(ctx, grad):
        x, conv_w, conv_b, bn_w, inds = ctx.saved_tensors
        rm, rv = ctx.stats
        g_w, g_b, g_gamma, g_beta = _pfn_backward(x.contiguous(), conv_w, conv_b, bn_w, rm, rv, ctx.training, ctx.eps,
                                                  grad, inds if ctx.canvas else None, ctx.hw[0], ctx.hw[1])
        return None, g_w, g_b, g_gamma, g_beta, None, None


class _ScatterFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, inds, module):
        ctx.save_for_backward(inds)
        ctx.P = x.shape[2]
        return module._run(x, inds)

    @staticmethod
    def backward(ctx, grad):
        L = _lib.load()
        (inds,) = ctx.saved_tensors
        grad = grad.contiguous()
        B, C, H, W = grad.shape
        out = torch.empty((B, C, ctx.P), dtype=torch.float32, device=grad.device)
        with torch.cuda.device(grad.device):
            rc = L.pp_scatter_backward(grad.data_ptr(), inds.data_ptr(), B, C, ctx.P, H, W, out.data_ptr(),
                                       _runtime.stream_ptr(grad.device))
        _lib.check(rc, "pp_scatter_backward")
        return out, None, None


def _wants_grad(x, module):
    return torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in module.parameters()))


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
        if _wants_grad(x, self):
            return _PFNFunction.apply(x, self.conv1.weight, self.conv1.bias, self.bn1.weight, self.bn1.bias, self, None)
        return self._run(x)


class PPScatter(nn.Module):
    """model/model.py:42-62: [B,C,P] + inds [B,P,3] (int64: flag, x, y) -> canvas [B,C,H,W]."""

    def __init__(self, device=None, canvas_height=None, canvas_width=None):
        super(PPScatter, self).__init__()
        self.device = device
        self.canvas_height = int(_cfg.canvas_height if canvas_height is None else canvas_height)
        self.canvas_width = int(_cfg.canvas_width if canvas_width is None else canvas_width)

    def forward(self, x, inds):
        if torch.is_grad_enabled() and isinstance(x, torch.Tensor) and x.requires_grad:
            return _ScatterFunction.apply(x, inds.contiguous(), self)
        return self._run(x, inds)

    def _run(self, x, inds):
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
        _runtime.poll_status(dev, "PPScatter (the reference raises IndexError on an index outside the canvas, model/model.py:61)")
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

    def forward(self, x, inds, return_features=False, out=None):
        # ``out`` / ``return_features`` are the streaming pipeline's arguments: that call is never differentiated
        if _wants_grad(x, self) and out is None and not return_features:
            _runtime.require_cuda(inds, "inds")
            return _PFNFunction.apply(x, self.conv1.weight, self.conv1.bias, self.bn1.weight, self.bn1.bias, self,
                                      inds.contiguous())
        return self._run(x, inds, return_features, out)

    @torch.no_grad()
    def _run(self, x, inds, return_features=False, out=None):
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
        _runtime.poll_status(dev, "PPFeatureScatter")
        return (canvas, feat) if return_features else canvas
