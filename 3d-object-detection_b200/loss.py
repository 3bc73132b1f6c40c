"""Host mirror of the reference's loss module (model/loss.py): ``PPLoss(b_ort, b_reg, b_cls, gamma, device)``
with ``forward(cls_tensor, reg_tensor, cls_targets, reg_targets) -> (p, cls_loss, reg_loss, ort_loss,
total_loss)``.  One fused CUDA pass computes the losses and the gradient of ``total_loss`` with respect to
both network outputs (pp_loss, csrc/loss.cu); ``total_loss.backward()`` hands those gradients to autograd.
The three component losses are returned detached (train.py:145-159 only prints them)."""
import torch
import torch.nn as nn

from . import _lib, _runtime
from .box_utils import Positives
from .config import PPConfig

_cfg = PPConfig()


class _PPLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cls_tensor, reg_tensor, cls_targets, reg_targets, gamma, alpha_pos, b_cls, b_reg, b_ort,
                num_classes):
        L = _lib.load()
        for t, n in ((cls_tensor, "cls_tensor"), (reg_tensor, "reg_tensor"), (cls_targets, "cls_targets"),
                     (reg_targets, "reg_targets")):
            if not isinstance(t, Positives) and not (t is None and isinstance(cls_targets, Positives)):
                _runtime.require_cuda(t, n)
        if cls_tensor.dtype != torch.float32 or reg_tensor.dtype != torch.float32:
            raise _lib.PPError("PPLoss: network outputs must be float32")
        if not reg_tensor.is_contiguous():
            raise _lib.PPError("PPLoss: reg_tensor must be a contiguous NCHW tensor (it is updated in place, "
                               "model/loss.py:50)")
        cls_c = cls_tensor.contiguous()
        B, CK, H, W = cls_c.shape
        K = int(num_classes)
        Ad = CK // K
        R = reg_tensor.shape[1] // Ad
        A = H * W * Ad
        pos = cls_targets if isinstance(cls_targets, Positives) else None
        if pos is not None:
            if (pos.n_sweeps, pos.n_anchors) != (B, A):
                raise _lib.PPError("PPLoss: the positives list was built for %d sweeps x %d anchors" % (pos.n_sweeps, pos.n_anchors))
        else:
            ct = cls_targets.to(torch.float32).contiguous().view(B, A, K)
            rt = reg_targets.to(torch.float32).contiguous().view(B, A, 9)
        dev = cls_c.device
        scores = torch.empty((B, A * K), dtype=torch.float32, device=dev)
        g_cls = torch.empty_like(cls_c)
        g_reg = torch.empty_like(reg_tensor)
        losses = torch.empty(4, dtype=torch.float32, device=dev)
        ws = _runtime.workspace(L.pp_loss_workspace_bytes(B, H, W, Ad), dev, "loss")
        with torch.cuda.device(dev):
            tail = (B, H, W, Ad, K, R, float(gamma), float(alpha_pos), float(b_cls), float(b_reg), float(b_ort),
                    scores.data_ptr(), g_cls.data_ptr(), g_reg.data_ptr(), losses.data_ptr(), ws.data_ptr(), ws.numel(),
                    _runtime.stream_ptr(dev))
            if pos is not None:
                rc = L.pp_loss_list(cls_c.data_ptr(), reg_tensor.data_ptr(), pos.anchor.data_ptr(), pos.cls.data_ptr(),
                                    pos.reg.data_ptr(), pos.offsets.data_ptr(), *tail)
            else:
                rc = L.pp_loss(cls_c.data_ptr(), reg_tensor.data_ptr(), ct.data_ptr(), rt.data_ptr(), *tail)
        _lib.check(rc, "pp_loss_list" if pos is not None else "pp_loss")
        ctx.mark_dirty(reg_tensor)
        ctx.applied = None
        ctx.save_for_backward(g_cls, g_reg)
        ctx.mark_non_differentiable(scores)
        return losses[3], scores, losses[0].detach(), losses[1].detach(), losses[2].detach(), reg_tensor

    @staticmethod
    def backward(ctx, g_total, g_scores, g_c, g_r, g_o, g_regout):
        g_cls, g_reg = ctx.saved_tensors
        if ctx.applied is not None:
            # a second backward through the same graph: out of place, the first result may be in use
            r = g_total.to(torch.float32) / ctx.applied
            return g_cls * r, g_reg * r, None, None, None, None, None, None, None, None
        # first (normally only) backward: scale in place; the kernel exits at once for an upstream gradient of 1
        g = g_total.detach().to(torch.float32).contiguous()
        with torch.cuda.device(g_cls.device):
            rc = _lib.load().pp_loss_scale_grads(g_cls.data_ptr(), g_cls.numel(), g_reg.data_ptr(), g_reg.numel(),
                                                 g.data_ptr(), None, _runtime.stream_ptr(g_cls.device))
        _lib.check(rc, "pp_loss_scale_grads")
        ctx.applied = g
        return g_cls, g_reg, None, None, None, None, None, None, None, None


class PPLoss(nn.Module):
    """model/loss.py:11-63.  ``cls_weights`` of the reference are unused there (dead)."""

    def __init__(self, b_ort, b_reg, b_cls, gamma, device=None, num_classes=None, alpha_pos=25.0):
        super(PPLoss, self).__init__()
        self.b_ort, self.b_reg, self.b_cls, self.gamma, self.device = b_ort, b_reg, b_cls, gamma, device
        self.num_classes = int(_cfg.num_classes if num_classes is None else num_classes)
        self.alpha_pos = float(alpha_pos)            # torch.Tensor([25]) in model/loss.py:41

    def forward(self, cls_tensor, reg_tensor, cls_targets, reg_targets=None):
        """``cls_targets`` may be a ``box_utils.Positives`` list (then ``reg_targets`` is not used): the targets are
        taken as zero everywhere except at the listed anchors (pp_loss_list)."""
        total, p, c_loss, r_loss, o_loss, _ = _PPLossFn.apply(
            cls_tensor, reg_tensor, cls_targets, reg_targets, self.gamma, self.alpha_pos, self.b_cls, self.b_reg,
            self.b_ort, self.num_classes)
        return p, c_loss, r_loss, o_loss, total
