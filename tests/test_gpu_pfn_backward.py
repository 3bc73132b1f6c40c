"""PFN / scatter backward (pp_pfn_backward, pp_scatter_backward through the autograd of the host mirrors) vs
the golden fixture from the reference's own modules + torch autograd, vs the fp64 oracle, and vs a float64
torch autograd run of the same layers at a larger size.  Tolerance: 1e-5 of each gradient tensor's largest
magnitude (fp32 arithmetic against fp64; sums of up to 2e6 terms)."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import GOLDEN

pytestmark = pytest.mark.gpu

KEYS = ("grad_weight", "grad_bias", "grad_bn_weight", "grad_bn_bias")


def _close(got, want, tol=1e-5):
    got = np.asarray(got, np.float64); want = np.asarray(want, np.float64)
    scale = max(np.abs(want).max(), 1e-30)
    assert np.abs(got - want).max() <= tol * scale, "max diff %g of scale %g" % (np.abs(got - want).max(), scale)


def _net(sd, cls=None):
    from pp_b200 import model as pm
    net = (cls or pm.PPFeatureNet)(9, 64).cuda()
    net.load_state_dict({k: torch.as_tensor(np.asarray(v)) for k, v in sd.items()})
    return net


def _grads(net):
    return {"grad_weight": net.conv1.weight.grad.cpu().numpy().reshape(64, 9), "grad_bias": net.conv1.bias.grad.cpu().numpy(),
            "grad_bn_weight": net.bn1.weight.grad.cpu().numpy(), "grad_bn_bias": net.bn1.bias.grad.cpu().numpy()}


@pytest.mark.parametrize("tag", ["pos", "mixed"])
@pytest.mark.parametrize("mode", ["train", "eval"])
def test_golden_from_reference_autograd(tag, mode):
    g = np.load(os.path.join(GOLDEN, "pfn_backward_small.npz"))
    sd = {k[len(tag) + 4:]: g[k] for k in g.files if k.startswith(tag + "/sd/")}
    net = _net(sd)
    net.train(mode == "train")
    y = net(torch.tensor(g["x"], dtype=torch.float32, device="cuda"))
    y.backward(torch.tensor(g[tag + "/g_feat"], dtype=torch.float32, device="cuda"))
    got = _grads(net)
    for k in KEYS:
        _close(got[k], g["%s/%s/%s" % (tag, mode, k)])


def test_golden_end_to_end_through_the_canvas():
    """PPFeatureNet -> PPScatter (two modules) and PPFeatureScatter (fused) against the reference's autograd."""
    from pp_b200 import model as pm
    g = np.load(os.path.join(GOLDEN, "pfn_backward_small.npz"))
    sd = {k[7:]: g[k] for k in g.files if k.startswith("pos/sd/")}
    gc = torch.zeros((2, 64, 600, 600), device="cuda")
    i = torch.tensor(g["e2e/g_canvas_index"].astype(np.int64), device="cuda")
    gc[i[:, 0], i[:, 1], i[:, 2], i[:, 3]] = torch.tensor(g["e2e/g_canvas_value"], dtype=torch.float32, device="cuda")
    x = torch.tensor(g["x"], dtype=torch.float32, device="cuda")
    inds = torch.tensor(g["inds"], device="cuda")
    net = _net(sd).train()
    canvas = pm.PPScatter(torch.device("cuda"))(net(x), inds)
    (canvas * gc).sum().backward()
    two = _grads(net)
    fused = _net(sd, pm.PPFeatureScatter).train()
    (fused(x, inds) * gc).sum().backward()
    one = _grads(fused)
    for k in KEYS:
        _close(two[k], g["e2e/" + k])
        _close(one[k], g["e2e/" + k])


@pytest.mark.parametrize("training", [True, False])
@pytest.mark.parametrize("N", [200, 30])
def test_against_float64_autograd_and_determinism(training, N):
    """A larger case (B=3, P=1500): most pillars are padding rows that repeat across sweeps, like the real
    tensors; N=30 takes the unaligned staging path.  The comparison run is plain torch in float64."""
    from pp_b200 import model as pm
    torch.manual_seed(4)
    B, P = 3, 1500
    x = torch.randn((B, 9, P, N), device="cuda") * 0.8
    x[:, :, 600:, :] = torch.randn((1, 9, P - 600, N), device="cuda") * 0.1
    x[:, :, :600, N // 2:] = x[:1, :, 600:1200, N // 2:]
    net = pm.PPFeatureNet(9, 64).cuda()
    with torch.no_grad():
        net.bn1.weight.copy_(torch.rand(64, device="cuda") + 0.5)
        net.bn1.weight[::7] *= -1
        net.bn1.weight[5] = 0.0
        net.bn1.running_mean.normal_(0.3, 0.2); net.bn1.running_var.uniform_(0.5, 2.0)
    net.train(training)
    g_out = torch.randn((B, 64, P), device="cuda")
    rm, rv = net.bn1.running_mean.double().clone(), net.bn1.running_var.double().clone()
    grads = []
    for _ in range(2):
        net.zero_grad()
        net(x).backward(g_out)
        grads.append(_grads(net))
    for k in KEYS:
        assert np.array_equal(grads[0][k], grads[1][k])                          # run-to-run deterministic
    prm = [p.detach().double().requires_grad_(True) for p in (net.conv1.weight, net.conv1.bias, net.bn1.weight, net.bn1.bias)]
    y = F.batch_norm(F.relu(F.conv2d(x.double(), prm[0], prm[1])), rm, rv, prm[2], prm[3], training, 0.1, net.bn1.eps)
    y.max(dim=3)[0].backward(g_out.double())
    want = {"grad_weight": prm[0].grad.cpu().numpy().reshape(64, 9), "grad_bias": prm[1].grad.cpu().numpy(),
            "grad_bn_weight": prm[2].grad.cpu().numpy(), "grad_bn_bias": prm[3].grad.cpu().numpy()}
    for k in KEYS:
        _close(grads[0][k], want[k])


def test_scatter_backward_bit_exact_and_input_gradient_refused():
    from pp_b200 import model as pm, _lib
    from oracle import pfn_backward as ob
    rng = np.random.default_rng(9)
    B, C, P, H, W = 2, 64, 300, 600, 600
    inds = np.zeros((B, P, 3), np.int64)
    for b in range(B):
        cells = rng.choice(H * W, 250, replace=False)
        inds[b, :250, 0] = 1; inds[b, :250, 1] = cells % W; inds[b, :250, 2] = cells // W
    inds[0, 7] = inds[0, 3]                                                       # a duplicated cell: both rows read it
    feat = torch.randn((B, C, P), device="cuda", requires_grad=True)
    gc = torch.randn((B, C, H, W), device="cuda")
    pm.PPScatter(torch.device("cuda"))(feat, torch.tensor(inds, device="cuda")).backward(gc)
    want = ob.scatter_backward(gc.cpu().numpy(), inds)
    assert np.array_equal(feat.grad.cpu().numpy().astype(np.float64), want)
    net = pm.PPFeatureNet(9, 64).cuda()
    with pytest.raises(_lib.PPError):
        net(torch.randn((1, 9, 8, 16), device="cuda", requires_grad=True))


def test_pillar_count_not_a_multiple_of_the_group_size():
    """B*P = 50: the last four-pillar group of the kernel is half empty."""
    from pp_b200 import model as pm
    torch.manual_seed(8)
    x = torch.randn((1, 9, 50, 24), device="cuda")
    net = pm.PPFeatureNet(9, 64).cuda().train()
    g_out = torch.randn((1, 64, 50), device="cuda")
    net(x).backward(g_out)
    got = _grads(net)
    prm = [p.detach().double().requires_grad_(True) for p in (net.conv1.weight, net.conv1.bias, net.bn1.weight, net.bn1.bias)]
    y = F.batch_norm(F.relu(F.conv2d(x.double(), prm[0], prm[1])), None, None, prm[2], prm[3], True, 0.1, net.bn1.eps)
    y.max(dim=3)[0].backward(g_out.double())
    want = {"grad_weight": prm[0].grad.cpu().numpy().reshape(64, 9), "grad_bias": prm[1].grad.cpu().numpy(),
            "grad_bn_weight": prm[2].grad.cpu().numpy(), "grad_bn_bias": prm[3].grad.cpu().numpy()}
    for k in KEYS:
        _close(got[k], want[k])
