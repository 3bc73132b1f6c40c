"""ORACLE (test infrastructure only): fp64 restatement of the reference's PPLoss forward
(model/loss.py:24-63) and of the gradient of its total loss with respect to the two network outputs.

    cls_tensor [B, A_d*K, H, W], reg_tensor [B, A_d*R, H, W]  (NCHW network outputs, A_d = anchors
    per cell = 6, K = 9 classes, R = cfg.DATA.REG_DIMS = 8: dx,dy,dz,dw,dl,dh,dt,ort), cls_targets
    [B, A, K], reg_targets [B, A, 9] = [positive flag, 7 regression targets, orientation bit] with
    A = H*W*A_d (utils/box_utils.py:70-109).

Pinned against the reference module itself by tests/golden/loss_small.npz (make_golden.py).
"""
import numpy as np


def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def pp_loss(cls_tensor, reg_tensor, cls_targets, reg_targets, b_ort, b_reg, b_cls, gamma, reg_dims=8,
            alpha_pos=25.0):
    """Returns dict(p, cls_loss, reg_loss, ort_loss, total, grad_cls, grad_reg, reg_after) in float64.
    grad_* are d total / d (network output) in the NCHW layout of the inputs; reg_after is the
    regression output after the reference's in-place tanh (model/loss.py:50)."""
    cls_tensor = np.asarray(cls_tensor, dtype=np.float64)
    reg_tensor = np.asarray(reg_tensor, dtype=np.float64)
    cls_targets = np.asarray(cls_targets, dtype=np.float64)
    reg_targets = np.asarray(reg_targets, dtype=np.float64)
    B = cls_tensor.shape[0]
    # model/loss.py:32-37  permute(0,2,3,1), flatten
    x = cls_tensor.transpose(0, 2, 3, 1).reshape(B, -1)
    t = cls_targets.reshape(B, -1)
    # :39-44  focal weights (detached)
    p = _sigmoid(x)
    pt = np.where(t == 1, p, 1 - p)
    at = np.where(t == 1, alpha_pos, 1.0)
    w = at * (1 - pt) ** gamma
    # :46  F.binary_cross_entropy_with_logits(x, t, weight=w), mean over all elements
    bce = np.maximum(x, 0) - x * t + np.log1p(np.exp(-np.abs(x)))
    n_cls = x.size
    cls_loss = float((w * bce).sum() / n_cls)
    g_x = w * (p - t) / n_cls
    grad_cls = g_x.reshape(cls_tensor.transpose(0, 2, 3, 1).shape).transpose(0, 3, 1, 2)

    # :48-53  permute to [B,H,W,A_d*R]; the in-place tanh of :50 indexes that LAST axis with 6, i.e. it
    # touches network channel 6 only (anchor 0's dt), not element 6 of every anchor; then reshape to [B, A, R]
    rp = reg_tensor.transpose(0, 2, 3, 1).copy()
    pre6 = rp[..., 6].copy()                                 # [B,H,W]
    rp[..., 6] = np.tanh(pre6)
    reg_after = rp.transpose(0, 3, 1, 2)
    r = rp.reshape(B, -1, reg_dims)
    dtanh = np.ones_like(rp)
    dtanh[..., 6] = 1 - np.tanh(pre6) ** 2                   # chain rule through the tanh, channel 6 only
    dtanh = dtanh.reshape(B, -1, reg_dims)
    # :54-57  positives, smooth-L1 (beta = 1) mean over n_pos * 7
    pos = reg_targets[..., 0] == 1
    n_pos = int(pos.sum())
    g_r = np.zeros_like(r)
    if n_pos > 0:
        d = r[pos][:, :7] - reg_targets[pos][:, 1:8]
        ad = np.abs(d)
        reg_loss = float(np.where(ad < 1, 0.5 * d * d, ad - 0.5).sum() / (n_pos * 7))
        gd = np.where(ad < 1, d, np.sign(d)) / (n_pos * 7)
        # :59-61  orientation BCE on element 7 vs target element 8, mean over n_pos
        o = r[pos][:, 7]
        ot = reg_targets[pos][:, 8]
        ort_loss = float((np.maximum(o, 0) - o * ot + np.log1p(np.exp(-np.abs(o)))).sum() / n_pos)
        go = (_sigmoid(o) - ot) / n_pos
        gpos = np.zeros((n_pos, reg_dims))
        gpos[:, :7] = b_reg * gd
        gpos[:, 7] = b_ort * go
        g_r[pos] = gpos * dtanh[pos]
    else:
        reg_loss = float("nan")                              # torch: mean over an empty tensor
        ort_loss = float("nan")
    grad_reg = g_r.reshape(reg_tensor.transpose(0, 2, 3, 1).shape).transpose(0, 3, 1, 2)
    total = b_cls * cls_loss + b_reg * reg_loss + b_ort * ort_loss      # :63
    return {"p": p, "cls_loss": cls_loss, "reg_loss": reg_loss, "ort_loss": ort_loss, "total": total,
            "grad_cls": b_cls * grad_cls, "grad_reg": grad_reg, "reg_after": reg_after, "n_pos": n_pos}
