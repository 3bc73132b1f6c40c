"""Backward of the fused input path (pp_input_path_backward, sparse formulation) against the dense module
backward (pp_pfn_backward on the materialised x, itself pinned to the reference's autograd) and against float64
torch autograd of the reference's layers on that x.  Tolerance 1e-5 of each gradient tensor's largest magnitude."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

KEYS = ("conv1.weight", "conv1.bias", "bn1.weight", "bn1.bias")


def _close(got, want, tol=1e-5):
    got = got.double().cpu().numpy().ravel(); want = want.double().cpu().numpy().ravel()
    scale = max(np.abs(want).max(), 1e-30)
    assert np.abs(got - want).max() <= tol * scale, "max diff %g of scale %g" % (np.abs(got - want).max(), scale)


@pytest.mark.parametrize("training", [True, False])
@pytest.mark.parametrize("B", [2, 3])
def test_fused_backward_equals_dense_backward_and_float64_autograd(training, B):
    import pp_b200
    from pp_b200 import model as pm, pipeline, synth
    cfg = pp_b200.PPConfig(max_pillars=3002 if B == 3 else 3000, max_points_per_pillar=48)   # 3002: a half-empty last group
    P, N = cfg.max_pillars, cfg.max_points_per_pillar
    mean = synth.make_data_mean(P, N, seed=3, dense=True)
    prm = synth.make_pfn_params(5, flip_gamma=True)
    path = pipeline.InputPath(cfg, device=torch.device("cuda"), data_mean=mean, pfn_params=prm, training=training, fused=True)
    sweeps = [synth.make_sweep(20 + b)[:9000 + 4000 * b] for b in range(B)]
    pts = torch.tensor(np.concatenate(sweeps), device="cuda")
    offs = [0] + list(np.cumsum([len(s) for s in sweeps]))
    sd0 = {k: v.clone() for k, v in path.net.state_dict().items()}
    canvas, inds, npil = path.pillarize_encode_train(pts, offs)
    assert canvas.requires_grad and not inds.requires_grad
    g_canvas = torch.randn(canvas.shape, device="cuda", generator=torch.Generator("cuda").manual_seed(1))
    # a second forward before the backward must not disturb the saved state (the graph owns its workspace)
    other = path.pillarize_encode(torch.tensor(synth.make_sweep(99)[:5000], device="cuda"), [0, 5000])
    canvas.backward(g_canvas)
    got = {k: dict(path.net.named_parameters())[k].grad.clone() for k in KEYS}
    assert int(npil.min()) > 500 and int(npil.max()) <= P

    # dense reference: x materialised by the same library, module backward (pp_pfn_backward)
    dense = pipeline.InputPath(cfg, device=torch.device("cuda"), data_mean=mean, pfn_params=prm, training=training, fused=False)
    x, inds2, npil2 = dense.pillarize(pts, offs)
    assert torch.equal(inds, inds2) and torch.equal(npil, npil2)
    net = pm.PPFeatureScatter(9, 64, cfg.canvas_height, cfg.canvas_width).cuda()
    net.load_state_dict(sd0)
    net.train(training)
    c2 = net(x, inds2)
    c2.backward(g_canvas)
    for k in KEYS:
        _close(got[k], dict(net.named_parameters())[k].grad)
    assert (canvas.detach() - c2.detach()).abs().max() <= 1e-5 * c2.abs().max() + 2e-4

    # float64 torch autograd of the reference's layers on the same x
    prm64 = [sd0[k].detach().double().requires_grad_(True) for k in KEYS]
    y = F.batch_norm(F.relu(F.conv2d(x.double(), prm64[0], prm64[1])), sd0["bn1.running_mean"].double().clone(),
                     sd0["bn1.running_var"].double().clone(), prm64[2], prm64[3], training, 0.1, 1e-5).max(dim=3)[0]
    out = torch.zeros(canvas.shape, dtype=torch.float64, device="cuda")
    ne = torch.nonzero(inds[:, :, 0])
    bb, pp_ = ne[:, 0], ne[:, 1]
    out[bb, :, inds[bb, pp_][:, 2], inds[bb, pp_][:, 1]] = y[bb, :, pp_]
    (out * g_canvas.double()).sum().backward()
    for k, t in zip(KEYS, prm64):
        _close(got[k], t.grad)


def test_fused_backward_edge_cases():
    """More than eight sweeps, one of them empty; N = 16 so many pillars overflow the point cap (no padding slot
    competes there); P = 500 so the pillar cap binds.  Against the dense module backward on the materialised x."""
    import pp_b200
    from pp_b200 import model as pm, pipeline, synth
    cfg = pp_b200.PPConfig(max_pillars=500, max_points_per_pillar=16)
    P, N = cfg.max_pillars, cfg.max_points_per_pillar
    mean = synth.make_data_mean(P, N, seed=6, dense=True)
    prm = synth.make_pfn_params(2, flip_gamma=True)
    path = pipeline.InputPath(cfg, device=torch.device("cuda"), data_mean=mean, pfn_params=prm, training=True, fused=True)
    sweeps = [synth.make_sweep(60 + b)[:6000] for b in range(9)]
    sweeps[4] = sweeps[4][:0]                                                     # an empty sweep
    sweeps[7] = sweeps[7][:40]                                                    # a nearly empty one
    pts = torch.tensor(np.concatenate(sweeps), device="cuda")
    offs = [0] + list(np.cumsum([len(s) for s in sweeps]))
    sd0 = {k: v.clone() for k, v in path.net.state_dict().items()}
    canvas, inds, npil = path.pillarize_encode_train(pts, offs)
    assert npil.tolist()[4] == 0 and npil.max().item() == P
    g_canvas = torch.randn(canvas.shape, device="cuda", generator=torch.Generator("cuda").manual_seed(2))
    canvas.backward(g_canvas)
    got = {k: dict(path.net.named_parameters())[k].grad.clone() for k in KEYS}
    dense = pipeline.InputPath(cfg, device=torch.device("cuda"), data_mean=mean, pfn_params=prm, training=True, fused=False)
    x, inds2, _ = dense.pillarize(pts, offs)
    net = pm.PPFeatureScatter(9, 64, cfg.canvas_height, cfg.canvas_width).cuda().train()
    net.load_state_dict(sd0)
    net(x, inds2).backward(g_canvas)
    for k in KEYS:
        _close(got[k], dict(net.named_parameters())[k].grad)
