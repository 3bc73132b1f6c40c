"""Positives-list hand-off K3 -> loss (pp_assign_targets_list, pp_loss_list): the list equals the
non-zero rows of the dense targets bit for bit and in ascending order; the loss fed by the list equals the loss
fed by the dense tensors (same per-element arithmetic; sums within fp32 reassociation) and the fp64 oracle."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _targets(B, n_gt, cfg=None):
    import pp_b200
    from pp_b200 import pipeline, synth
    cfg = cfg or pp_b200.PPConfig()
    path = pipeline.InputPath(cfg, device=torch.device("cuda"))
    gts = [synth.make_gt(40 + b, n_gt) for b in range(B)]
    sweeps = [synth.make_sweep(b)[:1000] for b in range(B)]
    batch = path.pack_host_batch(sweeps, gts)
    _, gt_dev = path.upload(batch)
    dense = path.targets(gt_dev, batch["gt_offsets"])
    path.targets_as_list = True
    lst = path.targets(gt_dev, batch["gt_offsets"])
    torch.cuda.synchronize()
    return path, dense, lst


@pytest.mark.parametrize("B,n_gt", [(2, 60), (1, 0), (5, 100)])
def test_list_equals_the_nonzero_rows_of_the_dense_targets(B, n_gt):
    path, (cls, reg, top, counts), (pos, none, top2, counts2) = _targets(B, n_gt)
    assert none is None and torch.equal(top, top2) and torch.equal(counts, counts2)
    A = cls.shape[1]
    nz = torch.nonzero((cls.view(B * A, 9) != 0).any(1) | (reg.view(B * A, 9) != 0).any(1)).flatten()
    offs = pos.offsets.cpu().tolist()
    n = offs[-1]
    assert n == nz.numel() and (n > 0) == (n_gt > 0)
    assert torch.equal(pos.anchor[:n].long(), nz)                                   # ascending (sweep, anchor)
    assert torch.equal(pos.cls[:n], cls.view(B * A, 9)[nz]) and torch.equal(pos.reg[:n], reg.view(B * A, 9)[nz])
    for b in range(B + 1):
        assert offs[b] == int((nz < b * A).sum())
    dc, dr = pos.dense()
    assert torch.equal(dc, cls) and torch.equal(dr, reg)


def test_list_overflow_is_reported():
    import pp_b200
    from pp_b200 import _lib, _runtime, box_utils, pipeline, synth
    path = pipeline.InputPath(device=torch.device("cuda"))
    gts = [synth.make_gt(1, 80)]
    batch = path.pack_host_batch([synth.make_sweep(0)[:100]], gts)
    _, g = path.upload(batch)
    a = path.ensure_anchors()
    full, _, _, _ = box_utils.assign_targets(a, g["corners"], g["centers"], g["wlh"], g["yaw"], g["cls"], batch["gt_offsets"],
                                             as_list=True)
    n_full = int(full.offsets[-1].item())
    assert n_full > 10
    pos, _, _, _ = box_utils.assign_targets(a, g["corners"], g["centers"], g["wlh"], g["yaw"], g["cls"], batch["gt_offsets"],
                                            as_list=True, capacity=10)
    # offsets are clamped to the capacity (the consumers index the list arrays up to offsets[B]); the kept rows are
    # the first ones in (sweep, anchor) order
    assert pos.offsets.tolist() == [0, 10]
    assert torch.equal(pos.anchor[:10], full.anchor[:10]) and torch.equal(pos.cls[:10], full.cls[:10])
    # the loss over the truncated list stays inside the arrays (run under compute-sanitizer by scripts/sanitize.sh)
    from pp_b200.loss import PPLoss
    cls = torch.randn((1, 54, 300, 300), device="cuda") - 3.0
    reg = torch.randn((1, 48, 300, 300), device="cuda")
    out = PPLoss(0.4, 1.0, 250.0, 2, torch.device("cuda"))(cls, reg, pos)
    assert torch.isfinite(out[4])
    with pytest.raises(_lib.PPError):
        _runtime.check_status(torch.device("cuda"), "targets")
    with pytest.raises(_lib.PPError):                      # Positives.dense() refuses a truncated list
        pos2, _, _, _ = box_utils.assign_targets(a, g["corners"], g["centers"], g["wlh"], g["yaw"], g["cls"],
                                                 batch["gt_offsets"], as_list=True, capacity=10)
        pos2.dense()


@pytest.mark.parametrize("B,n_gt", [(2, 60), (1, 0)])
def test_loss_from_the_list_equals_loss_from_dense_targets_and_oracle(B, n_gt):
    from oracle import loss as ol
    from pp_b200.loss import PPLoss
    path, (cls_t, reg_t, _, _), (pos, _, _, _) = _targets(B, n_gt)
    torch.manual_seed(B)
    cls = torch.randn((B, 54, 300, 300), device="cuda") * 1.5 - 3.0
    reg = torch.randn((B, 48, 300, 300), device="cuda")
    lossm = PPLoss(0.4, 1.0, 250.0, 2, torch.device("cuda"))

    def run(*targets):
        c = cls.clone().requires_grad_(True); r = reg.clone().requires_grad_(True)
        r2 = r * 1.0
        p, cl, rl, ol_, tot = lossm(c, r2, *targets)
        tot.backward()
        return p.detach(), torch.stack([cl, rl, ol_, tot]).detach(), c.grad, r.grad, r2.detach()

    a = run(cls_t, reg_t)
    b = run(pos)
    b2 = run(pos)
    assert torch.equal(a[0], b[0]) and torch.equal(a[2], b[2]) and torch.equal(a[3], b[3]) and torch.equal(a[4], b[4])
    if n_gt:
        assert torch.allclose(a[1], b[1], rtol=2e-6, atol=0)
    else:
        assert torch.isnan(b[1][1]) and torch.isnan(a[1][1]) and torch.allclose(a[1][0], b[1][0], rtol=2e-6)
    assert all(torch.equal(x, y) or (torch.isnan(x).any() and torch.equal(torch.isnan(x), torch.isnan(y))) for x, y in zip(b, b2))
    if B == 2:
        want = ol.pp_loss(cls.cpu().numpy(), reg.cpu().numpy(), cls_t.cpu().numpy(), reg_t.cpu().numpy(), 0.4, 1.0, 250.0, 2)
        got = b[1].cpu().numpy().astype(np.float64)
        ref = np.array([want["cls_loss"], want["reg_loss"], want["ort_loss"], want["total"]])
        assert np.abs(got - ref).max() <= 5e-6 * np.abs(ref).max()
        assert np.abs(b[2].cpu().numpy() - want["grad_cls"]).max() <= 1e-5 * np.abs(want["grad_cls"]).max()


def test_loss_list_with_more_positives_than_fit_in_shared_memory():
    """More than 8192 listed anchors: the ids are searched in global memory and the class rows read from there;
    also rows with a class value that is neither 0 nor 1 (general floats)."""
    from pp_b200.box_utils import Positives
    from pp_b200.loss import PPLoss
    rng = np.random.default_rng(12)
    B, H, W = 2, 300, 300
    A = H * W * 6
    n_per = 5000
    cls_t = torch.zeros((B, A, 9), device="cuda"); reg_t = torch.zeros((B, A, 9), device="cuda")
    for b in range(B):
        idx = torch.tensor(np.sort(rng.choice(A, n_per, replace=False)), device="cuda")
        cls_t[b, idx, torch.tensor(rng.integers(0, 9, n_per), device="cuda")] = 1
        reg_t[b, idx, 0] = 1
        reg_t[b, idx, 1:8] = torch.tensor(rng.normal(0, 1, (n_per, 7)), dtype=torch.float32, device="cuda")
        reg_t[b, idx, 8] = torch.tensor(rng.integers(0, 2, n_per), dtype=torch.float32, device="cuda")
    cls_t[0, 17, 3] = 0.25                                                        # a soft label
    reg_t[0, 17, 0] = 1
    nz = torch.nonzero((cls_t.view(B * A, 9) != 0).any(1) | (reg_t.view(B * A, 9) != 0).any(1)).flatten()
    assert nz.numel() > 8192
    offs = torch.tensor([int((nz < b * A).sum()) for b in range(B + 1)], dtype=torch.int32, device="cuda")
    pos = Positives(nz.int().contiguous(), cls_t.view(B * A, 9)[nz].contiguous(), reg_t.view(B * A, 9)[nz].contiguous(), offs, B, A)
    torch.manual_seed(3)
    cls = torch.randn((B, 54, H, W), device="cuda") * 1.5 - 3.0
    reg = torch.randn((B, 48, H, W), device="cuda")
    lossm = PPLoss(0.4, 1.0, 250.0, 2, torch.device("cuda"))

    def run(*targets):
        c = cls.clone().requires_grad_(True); r = reg.clone().requires_grad_(True)
        p, cl, rl, ol_, tot = lossm(c, r * 1.0, *targets)
        tot.backward()
        return p.detach(), torch.stack([cl, rl, ol_, tot]).detach(), c.grad, r.grad

    a, b = run(cls_t, reg_t), run(pos)
    assert torch.equal(a[0], b[0]) and torch.equal(a[2], b[2]) and torch.equal(a[3], b[3])
    assert torch.allclose(a[1], b[1], rtol=2e-6, atol=0)
