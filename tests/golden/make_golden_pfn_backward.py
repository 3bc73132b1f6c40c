"""Golden fixture for the PFN / scatter backward from the REFERENCE's own modules (model/model.py
PPFeatureNet, PPScatter) run in float64 on the CPU with torch autograd.  Build container only:

    python tests/golden/make_golden_pfn_backward.py   ->  tests/golden/pfn_backward_small.npz
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import _easydict_standin, REF   # noqa: E402


def main():
    assert os.path.isdir(REF)
    _easydict_standin()
    sys.path.insert(0, REF)
    from model.model import PPFeatureNet, PPScatter              # the reference's own modules
    rng = np.random.default_rng(77)
    B, D, P, N, C = 2, 9, 48, 16, 64
    out = {}
    x = rng.normal(0, 1.0, (B, D, P, N))
    x[:, :, 40:, :] = rng.normal(0, 0.05, (1, D, 8, N))          # "empty" pillars: the same values in every sweep
    inds = np.zeros((B, P, 3), np.int64)
    cells = np.stack([rng.choice(600 * 600, 40, replace=False) for _ in range(B)])
    inds[:, :40, 0] = 1; inds[:, :40, 1] = cells % 600; inds[:, :40, 2] = cells // 600
    out["x"] = x; out["inds"] = inds
    for tag, flip in (("pos", False), ("mixed", True)):
        net = PPFeatureNet(D, C).double()
        with torch.no_grad():
            if flip:
                net.bn1.weight.copy_(torch.from_numpy(rng.uniform(0.5, 1.5, C) * rng.choice([-1.0, 1.0], C)))
                net.bn1.bias.copy_(torch.from_numpy(rng.normal(0, 0.2, C)))
                net.bn1.running_mean.copy_(torch.from_numpy(rng.normal(1, 0.5, C)))
                net.bn1.running_var.copy_(torch.from_numpy(rng.uniform(0.5, 3, C)))
            for prm in list(net.parameters()) + [net.bn1.running_mean, net.bn1.running_var]:
                prm.copy_(prm.float().double())                   # float32-representable parameters
        for k, v in net.state_dict().items():
            out["%s/sd/%s" % (tag, k)] = v.clone().numpy()
        g_feat = rng.normal(0, 1.0, (B, C, P))
        out["%s/g_feat" % tag] = g_feat
        for mode in ("eval", "train"):                      # eval first: the train forward advances the running statistics
            net.train(mode == "train")
            net.zero_grad()
            xt = torch.from_numpy(x).requires_grad_(True)
            y = net(xt)
            y.backward(torch.from_numpy(g_feat))
            out["%s/%s/grad_weight" % (tag, mode)] = net.conv1.weight.grad.numpy().reshape(C, D).copy()
            out["%s/%s/grad_bias" % (tag, mode)] = net.conv1.bias.grad.numpy().copy()
            out["%s/%s/grad_bn_weight" % (tag, mode)] = net.bn1.weight.grad.numpy().copy()
            out["%s/%s/grad_bn_bias" % (tag, mode)] = net.bn1.bias.grad.numpy().copy()
            out["%s/%s/grad_x" % (tag, mode)] = xt.grad.numpy().copy()
        if tag == "pos":
            # PFN -> scatter -> a fixed random linear functional of the canvas, end to end
            net.train(True); net.zero_grad()
            sc = PPScatter(torch.device("cpu"))
            g_canvas = np.zeros((B, C, 600, 600))
            for b in range(B):
                g_canvas[b][:, inds[b, :40, 2], inds[b, :40, 1]] = rng.normal(0, 1, (C, 40))
            g_canvas[:, :, 0, 0] = rng.normal(0, 1, (B, C))       # the cell the flag-0 rows point at: must not leak
            torch.set_default_dtype(torch.float64)
            try:
                canvas = sc(net(torch.from_numpy(x)), torch.from_numpy(inds))
            finally:
                torch.set_default_dtype(torch.float32)
            (canvas * torch.from_numpy(g_canvas)).sum().backward()
            nz = np.nonzero(g_canvas)
            out["e2e/g_canvas_index"] = np.stack(nz, 1).astype(np.int32)
            out["e2e/g_canvas_value"] = g_canvas[nz]
            out["e2e/grad_weight"] = net.conv1.weight.grad.numpy().reshape(C, D).copy()
            out["e2e/grad_bias"] = net.conv1.bias.grad.numpy().copy()
            out["e2e/grad_bn_weight"] = net.bn1.weight.grad.numpy().copy()
            out["e2e/grad_bn_bias"] = net.bn1.bias.grad.numpy().copy()
            for k, v in net.state_dict().items():
                out["e2e/sd/%s" % k] = v.clone().numpy()
    np.savez_compressed(os.path.join(HERE, "pfn_backward_small.npz"), **out)
    print("wrote pfn_backward_small.npz", {k: v.shape for k, v in out.items() if "grad_weight" in k})


if __name__ == "__main__":
    main()
