"""End-to-end training steps through the drop-ins together: fused input path (pillarize_encode_train) -> a small
torch head standing in for the backbone / detection head -> PPLoss fed by K3's positives list -> backward into
conv1 / bn1 of the pillar encoder -> SGD.  Checks that every piece hands finite, non-trivial gradients to the
next and that a few steps on a fixed batch reduce the loss."""
import numpy as np
import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu


def test_three_sgd_steps_reduce_the_loss():
    import pp_b200
    from pp_b200 import pipeline, synth
    from pp_b200.loss import PPLoss
    torch.manual_seed(0)
    cfg = pp_b200.PPConfig()
    P, N = cfg.max_pillars, cfg.max_points_per_pillar
    path = pipeline.InputPath(cfg, device=torch.device("cuda"), data_mean=synth.make_data_mean(P, N, seed=0, dense=True),
                              pfn_params=synth.make_pfn_params(0), training=True, fused=True)
    path.targets_as_list = True
    sweeps = [synth.make_sweep(3), synth.make_sweep(4)]
    gts = [synth.make_gt(3, 40), synth.make_gt(4, 40)]
    batch = path.pack_host_batch(sweeps, gts)
    d_pts, gt_dev = path.upload(batch)
    pos, _, _, counts = path.targets(gt_dev, batch["gt_offsets"])
    assert int(pos.offsets[-1].item()) > 50
    head_c = nn.Conv2d(64, 54, 3, stride=2, padding=1).cuda()
    head_r = nn.Conv2d(64, 48, 3, stride=2, padding=1).cuda()
    with torch.no_grad():
        head_c.bias.fill_(-4.0)                                   # the usual focal-loss prior
    params = list(path.net.parameters()) + list(head_c.parameters()) + list(head_r.parameters())
    opt = torch.optim.SGD(params, lr=2e-3)
    lossm = PPLoss(0.2, 2.0, 1.0, 2, torch.device("cuda"))
    history = []
    for step in range(4):
        opt.zero_grad(set_to_none=True)
        canvas, inds, npil = path.pillarize_encode_train(d_pts, batch["offsets"])
        p, cl, rl, ol, total = lossm(head_c(canvas), head_r(canvas), pos)
        total.backward()
        for prm in path.net.parameters():
            assert prm.grad is not None and torch.isfinite(prm.grad).all() and prm.grad.abs().max() > 0
        opt.step()
        history.append(float(total.detach()))
    assert all(np.isfinite(history))
    assert history[-1] < history[0], history
