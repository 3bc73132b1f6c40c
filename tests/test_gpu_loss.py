"""Loss front-end (pp_loss through the PPLoss host mirror) vs the golden fixture from the reference's own
module and vs the fp64 oracle at the reference's full size.  Tolerance: 1e-5 relative (fp32 elementwise
arithmetic against fp64), with an absolute floor scaled to each tensor."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


def _close(a, b, rtol=1e-5, atol=0.0):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    err = np.abs(a - b) - (rtol * np.maximum(np.abs(a), np.abs(b)) + atol)
    assert err.max() <= 0, "max violation %g (max abs diff %g)" % (err.max(), np.abs(a - b).max())


def _run(cls, reg, cls_t, reg_t, b_ort, b_reg, b_cls, gamma):
    from pp_b200.loss import PPLoss
    loss = PPLoss(b_ort, b_reg, b_cls, gamma, torch.device("cuda"))
    ct = torch.tensor(cls, dtype=torch.float32, device="cuda", requires_grad=True)
    rt = torch.tensor(reg, dtype=torch.float32, device="cuda", requires_grad=True)
    ct2, rt2 = ct * 1.0, rt * 1.0                    # non-leaf, like a network output
    p, c, r, o, total = loss(ct2, rt2, torch.tensor(cls_t, device="cuda"), torch.tensor(reg_t, device="cuda"))
    total.backward()
    return p, c, r, o, total, ct.grad, rt.grad, rt2


@pytest.mark.parametrize("tag", ["cfg", "ort", "g3"])
def test_golden_from_reference_module(tag):
    g = np.load(os.path.join(GOLDEN, "loss_small.npz"))
    b_ort, b_reg, b_cls, gamma = [float(v) for v in g[tag + "/params"]]
    p, c, r, o, total, gc, gr, reg_after = _run(g[tag + "/cls"], g[tag + "/reg"], g[tag + "/cls_t"], g[tag + "/reg_t"],
                                                b_ort, b_reg, b_cls, gamma)
    want = g[tag + "/losses"]
    _close([float(c), float(r), float(o), float(total)], want, rtol=2e-6)
    _close(p.cpu().numpy(), g[tag + "/p"], atol=2e-7)
    _close(gc.cpu().numpy(), g[tag + "/grad_cls"], atol=1e-7 * np.abs(g[tag + "/grad_cls"]).max())
    _close(gr.cpu().numpy(), g[tag + "/grad_reg"], atol=1e-7 * np.abs(g[tag + "/grad_reg"]).max())
    _close(reg_after.detach().cpu().numpy(), g[tag + "/reg_after"], atol=2e-7)       # in-place tanh, channel 6 only


def test_full_size_against_oracle_and_determinism():
    """B=2, 300x300 feature map, 540000 anchors, targets from the assignment kernels' layout."""
    from oracle import loss as ol
    rng = np.random.default_rng(11)
    B, H, W = 2, 300, 300
    A = H * W * 6
    cls = rng.normal(-3.0, 1.5, (B, 54, H, W)).astype(np.float32)
    reg = rng.normal(0.0, 1.0, (B, 48, H, W)).astype(np.float32)
    cls_t = np.zeros((B, A, 9), np.float32); reg_t = np.zeros((B, A, 9), np.float32)
    for b in range(B):
        idx = rng.choice(A, 180, replace=False)
        idx[:6] = [0, 1, 5, A - 1, A - 6, 6 * W]                      # first / last anchors, channel-6 anchors (d = 0)
        cls_t[b, idx, rng.integers(0, 9, len(idx))] = 1
        reg_t[b, idx, 0] = 1
        reg_t[b, idx, 1:8] = rng.normal(0, 1.5, (len(idx), 7))
        reg_t[b, idx, 8] = rng.integers(0, 2, len(idx))
    want = ol.pp_loss(cls, reg, cls_t, reg_t, 0.3, 1.0, 250.0, 2)
    outs = [_run(cls, reg, cls_t, reg_t, 0.3, 1.0, 250.0, 2) for _ in range(2)]
    p, c, r, o, total, gc, gr, reg_after = outs[0]
    _close([float(c), float(r), float(o), float(total)], [want["cls_loss"], want["reg_loss"], want["ort_loss"], want["total"]],
           rtol=5e-6)
    _close(p.cpu().numpy(), want["p"], atol=2e-7)
    _close(gc.cpu().numpy(), want["grad_cls"], atol=1e-6 * np.abs(want["grad_cls"]).max())
    _close(gr.cpu().numpy(), want["grad_reg"], atol=1e-6 * np.abs(want["grad_reg"]).max())
    _close(reg_after.detach().cpu().numpy(), want["reg_after"], atol=2e-7)
    assert int((gr != 0).any(dim=1).sum()) <= 2 * 180
    for a, b2 in zip(outs[0][:7], outs[1][:7]):                     # run-to-run deterministic
        assert torch.equal(a, b2)


def test_no_positive_anchor_gives_nan_like_the_reference():
    rng = np.random.default_rng(3)
    cls = rng.normal(0, 1, (1, 54, 5, 7)).astype(np.float32); reg = rng.normal(0, 1, (1, 48, 5, 7)).astype(np.float32)
    z = np.zeros((1, 5 * 7 * 6, 9), np.float32)
    p, c, r, o, total, gc, gr, _ = _run(cls, reg, z, z, 0.0, 1.0, 250.0, 2)
    assert torch.isnan(r) and torch.isnan(total) and torch.isfinite(c)
    assert not gr.any() and torch.isfinite(gc).all()


@pytest.mark.parametrize("tma", [1, 0])
@pytest.mark.parametrize("hw", [(8, 36), (3, 44), (1, 4)])
def test_tile_edges_on_both_kernels(tma, hw):
    """H*W % 4 == 0 takes the TMA-staged kernel: full tiles, a ragged last tile and a single tiny tile; the
    generic tile kernel must agree with the same oracle."""
    from oracle import loss as ol
    from pp_b200 import _lib
    H, W = hw
    rng = np.random.default_rng(H * 100 + W)
    B, A = 3, H * W * 6
    cls = rng.normal(-1.0, 2.5, (B, 54, H, W)).astype(np.float32)
    reg = rng.normal(0.0, 1.2, (B, 48, H, W)).astype(np.float32)
    cls_t = np.zeros((B, A, 9), np.float32); reg_t = np.zeros((B, A, 9), np.float32)
    for b in range(B):
        idx = rng.choice(A, min(9, A), replace=False)
        cls_t[b, idx, rng.integers(0, 9, len(idx))] = 1
        reg_t[b, idx, 0] = 1
        reg_t[b, idx, 1:8] = rng.normal(0, 1.5, (len(idx), 7))
        reg_t[b, idx, 8] = rng.integers(0, 2, len(idx))
    reg_t[0, 1, 0] = 0.5                                               # a flag that is not exactly 1 is not a positive
    reg_t[0, 2, 1] = 1.0                                               # a 1 outside the flag column is not one either
    want = ol.pp_loss(cls, reg, cls_t, reg_t, 0.7, 2.0, 250.0, 2)
    L = _lib.load()
    L.pp_set_option(b"loss_tma", tma)
    try:
        p, c, r, o, total, gc, gr, reg_after = _run(cls, reg, cls_t, reg_t, 0.7, 2.0, 250.0, 2)
    finally:
        L.pp_set_option(b"loss_tma", 1)
    _close([float(c), float(r), float(o), float(total)], [want["cls_loss"], want["reg_loss"], want["ort_loss"], want["total"]],
           rtol=5e-6)
    _close(p.cpu().numpy(), want["p"], atol=2e-7)
    _close(gc.cpu().numpy(), want["grad_cls"], atol=1e-6 * np.abs(want["grad_cls"]).max())
    _close(gr.cpu().numpy(), want["grad_reg"], atol=1e-6 * np.abs(want["grad_reg"]).max())
    _close(reg_after.detach().cpu().numpy(), want["reg_after"], atol=2e-7)


def test_upstream_gradient_scale_and_second_backward():
    from pp_b200.loss import PPLoss
    g = np.load(os.path.join(GOLDEN, "loss_small.npz"))
    b_ort, b_reg, b_cls, gamma = [float(v) for v in g["ort/params"]]
    loss = PPLoss(b_ort, b_reg, b_cls, gamma, torch.device("cuda"))
    ct = torch.tensor(g["ort/cls"], dtype=torch.float32, device="cuda", requires_grad=True)
    rt = torch.tensor(g["ort/reg"], dtype=torch.float32, device="cuda", requires_grad=True)
    total = loss(ct * 1.0, rt * 1.0, torch.tensor(g["ort/cls_t"], device="cuda"), torch.tensor(g["ort/reg_t"], device="cuda"))[4]
    (total * 3.0).backward(retain_graph=True)                         # loss scaling, as AMP's GradScaler does
    _close(ct.grad.cpu().numpy(), 3.0 * g["ort/grad_cls"], atol=3e-7 * np.abs(g["ort/grad_cls"]).max())
    _close(rt.grad.cpu().numpy(), 3.0 * g["ort/grad_reg"], atol=3e-7 * np.abs(g["ort/grad_reg"]).max())
    (total * 0.5).backward()                                          # accumulates into .grad: 3.5 x
    _close(ct.grad.cpu().numpy(), 3.5 * g["ort/grad_cls"], atol=5e-7 * np.abs(g["ort/grad_cls"]).max())
    _close(rt.grad.cpu().numpy(), 3.5 * g["ort/grad_reg"], atol=5e-7 * np.abs(g["ort/grad_reg"]).max())
