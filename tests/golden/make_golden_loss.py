"""Golden fixture for the loss front-end from the REFERENCE's own model/loss.py PPLoss (CPU, float64),
including autograd gradients of the total loss.  Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_loss.py      ->  tests/golden/loss_small.npz
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import _easydict_standin, REF   # noqa: E402


def make_case(seed, B=2, H=6, W=5, n_pos=7):
    rng = np.random.default_rng(seed)
    A = H * W * 6
    cls = rng.normal(-2.0, 2.0, (B, 54, H, W))
    reg = rng.normal(0.0, 1.5, (B, 48, H, W))                  # 6 anchors x cfg.DATA.REG_DIMS = 8
    cls_t = np.zeros((B, A, 9)); reg_t = np.zeros((B, A, 9))
    for b in range(B):
        for a in rng.choice(A, n_pos, replace=False):
            cls_t[b, a, rng.integers(0, 9)] = 1
            if rng.random() < 0.3:
                cls_t[b, a, rng.integers(0, 9)] = 1           # two class bits on one row (forced matches)
            reg_t[b, a, 0] = 1
            reg_t[b, a, 1:8] = rng.normal(0, 1.2, 7)
            reg_t[b, a, 8] = float(rng.random() < 0.5)
    return cls, reg, cls_t, reg_t


def main():
    assert os.path.isdir(REF)
    _easydict_standin()
    sys.path.insert(0, REF)
    from model.loss import PPLoss                                # the reference's own module
    out = {}
    for tag, seed, (b_ort, b_reg, b_cls, gamma) in [("cfg", 1, (0, 1, 250, 2)), ("ort", 2, (0.5, 2.0, 10.0, 2)),
                                                    ("g3", 3, (1.0, 1.0, 1.0, 3))]:
        cls, reg, cls_t, reg_t = make_case(seed)
        loss = PPLoss(b_ort, b_reg, b_cls, gamma, torch.device("cpu"))
        ct = torch.tensor(cls, dtype=torch.float64, requires_grad=True)
        rt = torch.tensor(reg, dtype=torch.float64, requires_grad=True)
        # the reference writes tanh into a view of its input: give it a non-leaf tensor like the model output
        ct2, rt2 = ct * 1.0, rt * 1.0
        orig_tensor = torch.Tensor
        torch.Tensor = lambda v: torch.tensor(v, dtype=torch.float64)   # model/loss.py:41 builds float32 constants
        try:
            p, c_loss, r_loss, o_loss, total = loss(ct2, rt2, torch.tensor(cls_t), torch.tensor(reg_t))
        finally:
            torch.Tensor = orig_tensor
        total.backward()
        for k, v in [("cls", cls), ("reg", reg), ("cls_t", cls_t), ("reg_t", reg_t), ("p", p.detach().numpy()),
                     ("losses", np.array([float(c_loss), float(r_loss), float(o_loss), float(total)])),
                     ("grad_cls", ct.grad.numpy()), ("grad_reg", rt.grad.numpy()),
                     ("reg_after", rt2.detach().numpy()), ("params", np.array([b_ort, b_reg, b_cls, gamma], float))]:
            out["%s/%s" % (tag, k)] = v
    np.savez_compressed(os.path.join(HERE, "loss_small.npz"), **out)
    print("wrote loss_small.npz", {k: v.shape for k, v in out.items() if k.startswith("cfg/")})


if __name__ == "__main__":
    main()
