"""fp64 restatement of what torch autograd computes through /root/reference/model/model.py:31-40
(PPFeatureNet.forward: conv1 1x1 -> relu -> bn1 -> max over N) and :53-62 (PPScatter.forward), i.e. the
backward the reference's train.py:147 runs.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Written as the plain layer-by-layer chain rule over the dense [B,C,P,N] activations (numpy), independent of
the moment formulation the CUDA kernel uses.  Pinned by tests/golden/pfn_backward_small.npz, which
tests/golden/make_golden_pfn_backward.py generates from the reference's own modules with torch autograd."""
import numpy as np


def pfn_backward(x, weight, bias, bn_weight, grad_out, training, eps=1e-5, running_mean=None, running_var=None):
    """x [B,D,P,N], weight [C,D], grad_out [B,C,P] = dL/d(PPFeatureNet output).
    Returns dict(grad_weight [C,D], grad_bias [C], grad_bn_weight [C], grad_bn_bias [C], grad_x [B,D,P,N])."""
    x = np.asarray(x, np.float64); W = np.asarray(weight, np.float64).reshape(len(bias), -1)
    g = np.asarray(grad_out, np.float64); gamma = np.asarray(bn_weight, np.float64)
    z = np.einsum('cd,bdpn->bcpn', W, x) + np.asarray(bias, np.float64)[None, :, None, None]     # conv1 (:36)
    r = np.maximum(z, 0.0)                                                                       # relu (:37)
    if training:                                                                                 # bn1 (:38)
        mu = r.mean(axis=(0, 2, 3)); var = r.var(axis=(0, 2, 3))
    else:
        mu = np.asarray(running_mean, np.float64); var = np.asarray(running_var, np.float64)
    s = np.sqrt(var + eps)
    xhat = (r - mu[None, :, None, None]) / s[None, :, None, None]
    y = gamma[None, :, None, None] * xhat                                                        # (+ beta: no effect on argmax)
    idx = y.argmax(axis=3)                                                                       # torch.max(dim=3) (:39)
    dy = np.zeros_like(y)
    np.put_along_axis(dy, idx[..., None], g[..., None], axis=3)
    d_gamma = (dy * xhat).sum(axis=(0, 2, 3))
    d_beta = dy.sum(axis=(0, 2, 3))
    if training:
        m = float(x.shape[0] * x.shape[2] * x.shape[3])
        dr = (gamma / s)[None, :, None, None] * (dy - (d_beta / m)[None, :, None, None]
                                                 - xhat * (d_gamma / m)[None, :, None, None])
    else:
        dr = (gamma / s)[None, :, None, None] * dy
    dz = dr * (z > 0)
    return {"grad_weight": np.einsum('bcpn,bdpn->cd', dz, x), "grad_bias": dz.sum(axis=(0, 2, 3)),
            "grad_bn_weight": d_gamma, "grad_bn_bias": d_beta, "grad_x": np.einsum('cd,bcpn->bdpn', W, dz)}


def scatter_backward(grad_canvas, inds):
    """Backward of ``out[batch,:,y,x] = feat[batch,:,pillar]`` (model/model.py:61) as torch's index_put
    derivative defines it: every indexed row reads the canvas gradient at its cell, rows with
    inds[b,p,0] == 0 get zero.  grad_canvas [B,C,H,W], inds [B,P,3] -> [B,C,P]."""
    g = np.asarray(grad_canvas, np.float64)
    inds = np.asarray(inds)
    B, P = inds.shape[:2]
    out = np.zeros((B, g.shape[1], P))
    for b in range(B):
        sel = np.nonzero(inds[b, :, 0])[0]
        out[b][:, sel] = g[b][:, inds[b, sel, 2], inds[b, sel, 1]]
    return out
