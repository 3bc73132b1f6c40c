"""fp64 restatement of /root/reference/model/model.py:13-62 (PPFeatureNet, PPScatter).
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The reference modules themselves import fine in the build container (with an ``easydict``
stand-in) but /root/reference does not travel to the GPU box, so this restatement is what the
GPU tests compare against; tests/golden/pfn_*.npz (generated from the reference modules by
tests/golden/make_golden.py) pin it.
"""
import torch


def pfn_forward(x, weight, bias, bn_weight, bn_bias, running_mean, running_var, training,
                momentum=0.1, eps=1e-5):
    """model/model.py:36-39: conv1 (1x1) -> relu -> bn1 -> max over dim 3, in float64.

    x [B,D,P,N]; weight [C,D] (conv1.weight[:, :, 0, 0]); returns (out [B,C,P] f64,
    new_running_mean [C], new_running_var [C]) following nn.BatchNorm2d semantics: training uses
    the biased batch variance to normalise and folds the UNBIASED one into running_var."""
    x = x.double()
    W = weight.double().reshape(weight.shape[0], -1)
    y = torch.einsum('cd,bdpn->bcpn', W, x) + bias.double().view(1, -1, 1, 1)
    y = torch.relu(y)
    if training:
        n = y.numel() // y.shape[1]
        mean = y.mean(dim=(0, 2, 3))
        var = y.var(dim=(0, 2, 3), unbiased=False)
        new_rm = (1 - momentum) * running_mean.double() + momentum * mean
        new_rv = (1 - momentum) * running_var.double() + momentum * var * (n / max(n - 1, 1))
    else:
        mean, var = running_mean.double(), running_var.double()
        new_rm, new_rv = mean, var
    z = (y - mean.view(1, -1, 1, 1)) / torch.sqrt(var.view(1, -1, 1, 1) + eps)
    z = z * bn_weight.double().view(1, -1, 1, 1) + bn_bias.double().view(1, -1, 1, 1)
    return z.max(dim=3)[0], new_rm, new_rv


def scatter(x, inds, canvas_h, canvas_w):
    """model/model.py:53-62: zeros canvas; rows with inds[b,p,0] != 0 write x[b,:,p] to
    out[b,:,inds[b,p,2],inds[b,p,1]]."""
    out = torch.zeros(x.shape[0], x.shape[1], int(canvas_h), int(canvas_w), dtype=x.dtype)
    non_empty = torch.nonzero(inds[:, :, 0])
    batch = non_empty[:, 0]
    pillar = non_empty[:, 1]
    x_inds = inds[batch, pillar][:, 1]
    y_inds = inds[batch, pillar][:, 2]
    out[batch, :, y_inds, x_inds] = x[batch, :, pillar]
    return out


def reference_forward_f32(x, weight, bias, bn_weight, bn_bias, running_mean, running_var, training,
                          momentum=0.1, eps=1e-5):
    """model/model.py:36-39 with the same float32 library ops the reference module calls
    (conv2d 1x1 -> relu -> batch_norm -> max over dim 3).  Used as the timed CPU baseline of
    the PFN stage; running statistics are updated in place when training."""
    import torch.nn.functional as F
    y = F.conv2d(x, weight.reshape(weight.shape[0], -1, 1, 1), bias)
    y = F.relu(y)
    y = F.batch_norm(y, running_mean, running_var, bn_weight, bn_bias, training, momentum, eps)
    return torch.max(y, dim=3)[0]
