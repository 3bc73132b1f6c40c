"""Times the loss front-end (pp_loss: forward + both gradients in one pass) at the reference's training shape
and, beside it, the same arithmetic written as plain torch ops with autograd on the same GPU (what the reference's
model/loss.py launches).  Prints per-kernel times from the library's profiler and the algorithmic HBM bytes."""
import argparse
import json
import sys, os

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pp_b200
from pp_b200 import _lib
from pp_b200.loss import PPLoss


def eager(cls, reg, cls_t, reg_t, b_ort, b_reg, b_cls, gamma):
    B = cls.shape[0]
    x = cls.permute(0, 2, 3, 1).reshape(B, -1)
    t = cls_t.reshape(B, -1)
    p = torch.sigmoid(x)
    pos = t == 1
    w = (torch.where(pos, 25.0, 1.0) * (1 - torch.where(pos, p, 1 - p)) ** gamma).detach()
    lc = F.binary_cross_entropy_with_logits(x, t, weight=w)
    r = reg.permute(0, 2, 3, 1)
    r[..., 6] = torch.tanh(r[..., 6])
    r = r.reshape(B, -1, 8)
    sel = torch.where(reg_t[..., 0] == 1)
    lr = F.smooth_l1_loss(r[sel][..., :7], reg_t[sel][..., 1:8])
    lo = F.binary_cross_entropy_with_logits(r[sel][..., 7], reg_t[sel][..., 8])
    return p, b_cls * lc + b_reg * lr + b_ort * lo


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--iters", type=int, default=20)
    a = ap.parse_args()
    B, H, W = a.batch, 300, 300
    A = H * W * 6
    g = torch.Generator(device="cuda").manual_seed(5)
    cls0 = torch.randn((B, 54, H, W), device="cuda", generator=g) * 1.5 - 3
    reg0 = torch.randn((B, 48, H, W), device="cuda", generator=g)
    cls_t = torch.zeros((B, A, 9), device="cuda"); reg_t = torch.zeros((B, A, 9), device="cuda")
    rng = np.random.default_rng(0)
    for b in range(B):
        idx = torch.tensor(rng.choice(A, 200, replace=False), device="cuda")
        cls_t[b, idx, torch.tensor(rng.integers(0, 9, 200), device="cuda")] = 1
        reg_t[b, idx, 0] = 1
        reg_t[b, idx, 1:8] = torch.randn((200, 7), device="cuda", generator=g)
        reg_t[b, idx, 8] = 1
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    loss = PPLoss(0.3, 1.0, 250.0, 2, torch.device("cuda"))

    def ours():
        c = cls0.clone().requires_grad_(True); r = reg0.clone().requires_grad_(True)
        r2 = r * 1.0
        flush.zero_()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        p, _, _, _, tot = loss(c, r2, cls_t, reg_t)
        tot.backward(inputs=[c, r2])
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e3

    def theirs():
        c = cls0.clone().requires_grad_(True); r = reg0.clone().requires_grad_(True)
        r2 = r * 1.0
        flush.zero_()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        p, tot = eager(c, r2, cls_t, reg_t, 0.3, 1.0, 250.0, 2)
        tot.backward(inputs=[c, r2])
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e3

    # the same targets as a positives list (what pp_assign_targets_list emits)
    from pp_b200.box_utils import Positives
    nzi = torch.nonzero((cls_t.view(B * A, 9) != 0).any(1) | (reg_t.view(B * A, 9) != 0).any(1)).flatten()
    offs = torch.tensor([int((nzi < b * A).sum()) for b in range(B + 1)], dtype=torch.int32, device="cuda")
    pos = Positives(nzi.int().contiguous(), cls_t.view(B * A, 9)[nzi].contiguous(), reg_t.view(B * A, 9)[nzi].contiguous(),
                    offs, B, A)

    def ours_list():
        c = cls0.clone().requires_grad_(True); r = reg0.clone().requires_grad_(True)
        r2 = r * 1.0
        flush.zero_()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        tot = loss(c, r2, pos)[4]
        tot.backward(inputs=[c, r2])
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e3

    for _ in range(3):
        ours(); theirs(); ours_list()
    t_ours = float(np.median([ours() for _ in range(a.iters)]))
    t_eager = float(np.median([theirs() for _ in range(a.iters)]))
    t_list = float(np.median([ours_list() for _ in range(a.iters)]))
    L = _lib.load()
    L.pp_profile_enable(1)
    for _ in range(a.iters):
        ours(); ours_list()
    prof = {k: round(ms * 1e3 / n, 2) for k, (n, ms) in _lib.profile_report().items() if k.startswith("k_loss")}
    L.pp_profile_enable(0)
    nc, nr = B * 54 * H * W * 4, B * 48 * H * W * 4
    alg = {"k_loss_cls_tma": 4 * nc, "k_loss_cls_list": 3 * nc, "k_loss_reg": B * A * 9 * 4}
    print(json.dumps({"batch": B, "pp_loss_fwd_bwd_us": round(t_ours, 1), "pp_loss_list_fwd_bwd_us": round(t_list, 1), "torch_eager_fwd_bwd_us": round(t_eager, 1),
                      "kernels_us": prof, "alg_bytes": alg,
                      "GBps": {k: round(alg[k] / prof[k] / 1e3, 1) for k in alg if k in prof}}))


if __name__ == "__main__":
    main()
