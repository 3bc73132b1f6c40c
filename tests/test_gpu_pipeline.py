"""The whole path through the host-facing call (InputPath.step_host) vs the oracle, plus
idempotence / batch-order properties at BASELINE sizes."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_step_host_matches_oracle_composition():
    import pp_b200
    from oracle import config as ocfg, glue, pfn, targets as T
    from helpers import boxes_from_gt
    from pp_b200 import pipeline, synth
    cfg = pp_b200.PPConfig(max_pillars=2000, max_points_per_pillar=32, fm_height=60, fm_width=60)
    P, N = cfg.max_pillars, cfg.max_points_per_pillar
    mean = synth.make_data_mean(P, N, seed=2)
    prm = synth.make_pfn_params(3, flip_gamma=True)
    path = pipeline.InputPath(cfg, data_mean=mean, pfn_params=prm, training=True)
    sweeps = [synth.make_sweep(s)[:15000] for s in (1, 2)]
    gcfg = pp_b200.PPConfig(canvas_width=120, canvas_height=120)
    gts = [synth.make_gt(s, G, gcfg) for s, G in ((1, 9), (2, 14))]
    for g in gts:                       # keep the boxes on the 60x60 anchor lattice, flip uses H=600
        g["centers"][:, 1] = 599 - g["centers"][:, 1]
    batch = path.pack_host_batch(sweeps, gts)
    canvas, cls, reg, npil, counts = path.step_host(batch)
    torch.cuda.synchronize()
    # oracle composition
    xs, inds = [], []
    for s in sweeps:
        x, i = glue.pillarize(s.astype(np.float64), torch.from_numpy(mean), P, N)
        xs.append(x); inds.append(i)
    x = torch.stack(xs); ind = torch.stack(inds)
    t = lambda a: torch.from_numpy(a)
    y, rm, rv = pfn.pfn_forward(x, t(prm["conv_w"]), t(prm["conv_b"]), t(prm["bn_w"]), t(prm["bn_b"]),
                                t(prm["running_mean"]), t(prm["running_var"]), True)
    want = pfn.scatter(y, ind, cfg.canvas_height, cfg.canvas_width).numpy()
    got = canvas.cpu().numpy().astype(np.float64)
    # tolerance of tests/test_gpu_pfn.py: 1e-5 relative + 2e-6 + 1e-6 * amp (conditioning of the fp32 dot product)
    yy = torch.relu(torch.einsum('cd,bdpn->bcpn', t(prm["conv_w"]).double(), x.double()) + t(prm["conv_b"]).double().view(1, -1, 1, 1))
    sigma = torch.sqrt(yy.var(dim=(0, 2, 3), unbiased=False) + 1e-5)
    absdot = torch.einsum('cd,bdpn->bcpn', t(prm["conv_w"]).double().abs(), x.double().abs()).amax(dim=3) + t(prm["conv_b"]).double().abs().view(1, -1, 1)
    amp = float((absdot * (t(prm["bn_w"]).double().abs() / sigma).view(1, -1, 1)).max())
    assert (np.abs(got - want) <= 1e-5 * np.maximum(np.abs(got), np.abs(want)) + 2e-6 + 1e-6 * amp).all()
    assert np.array_equal(got != 0, want != 0)
    boxes, corners, centers, _ = T.make_anchor_boxes(60, 60)
    for b, gt in enumerate(gts):
        g = boxes_from_gt(gt, T.Box, ocfg.CLASS_NAMES)
        gc, gcor = T.boxes_to_image_space(g)
        c0, r0, ious = T.create_target(corners, gcor, centers, gc, boxes, g, return_ious=True)
        assert np.array_equal(cls[b].cpu().numpy(), c0.astype(np.float32))
        r = reg[b].cpu().numpy().astype(np.float64); r0 = r0.astype(np.float32).astype(np.float64)
        assert (np.abs(r - r0) <= 1e-5 * np.maximum(np.abs(r), np.abs(r0)) + 1e-7).all()
        # counters: anchors positive by threshold (strict >), per-GT forced matches kept (best anchor != 0)
        assert int(counts[b, 0]) == int((ious.max(1) > 0.6).sum())
        assert int(counts[b, 1]) == int((ious.argmax(0) != 0).sum())


def test_full_size_batch_properties():
    """BASELINE config 2/4 sizes (B=4, P=24000, N=200, C=64, 600x600 canvas, 540000 anchors):
    idempotence, batch-order independence of pillars/targets, canvas support == pillar cells."""
    from pp_b200 import pipeline, synth
    path = pipeline.InputPath(data_mean=synth.make_data_mean(24000, 200, dense=False),
                              pfn_params=synth.make_pfn_params(0), training=False)
    sweeps = [synth.make_sweep(s) for s in range(4)]
    gts = [synth.make_gt(s, 100) for s in range(4)]
    b1 = path.pack_host_batch(sweeps, gts)
    c1, cls1, reg1, n1, k1 = [t.clone() for t in path.step_host(b1)]
    c2, cls2, reg2, n2, k2 = path.step_host(b1)
    assert torch.equal(c1, c2) and torch.equal(cls1, cls2) and torch.equal(reg1, reg2)     # idempotent
    b3 = path.pack_host_batch(sweeps[::-1], gts[::-1])
    c3, cls3, reg3, n3, k3 = path.step_host(b3)
    assert torch.equal(cls3.flip(0), cls1) and torch.equal(reg3.flip(0), reg1)
    assert torch.equal(n3.flip(0), n1)
    assert torch.equal(c3.flip(0), c1)          # eval-mode BN: per-sweep result independent of the batch
    x, inds, npil = path.pillarize(b1["points"].cuda(), b1["offsets"])
    for b in range(4):
        n = int(npil[b]); ii = inds[b, :n].cpu().numpy()
        support = torch.zeros(600, 600, dtype=torch.bool)
        support[ii[:, 2], ii[:, 1]] = True
        nz = (c1[b] != 0).any(0).cpu()
        assert not (nz & ~support).any()
    assert int(k1[:, 0].sum()) > 50


def test_async_streaming_steps_equal_the_blocking_call():
    """step_host_async (copy stream, two staging buffers, pinned counters) over alternating batches
    gives exactly what step_host gives for each of them, also with two steps in flight."""
    from pp_b200 import pipeline, synth
    path = pipeline.InputPath(data_mean=synth.make_data_mean(24000, 200, dense=False),
                              pfn_params=synth.make_pfn_params(0), training=False)
    batches = []
    for k in range(3):
        sweeps = [synth.make_sweep(10 * k + s) for s in range(2)]
        gts = [synth.make_gt(10 * k + s, 20 + 7 * k) for s in range(2)]
        batches.append(path.pack_host_batch(sweeps, gts))
    want = []
    for b in batches:
        c, cls, reg, n, k = path.step_host(b)
        want.append((c.clone(), cls.clone(), reg.clone(), n.cpu(), k.cpu()))
    handles, got = [], []
    for i in [0, 1, 2, 0, 2, 1]:
        h = path.step_host_async(batches[i])            # fresh outputs per step (no shared `out`)
        assert h.canvas is not None
        handles.append((i, h))
        if len(handles) > 2:
            j, hj = handles.pop(0)
            got.append((j, hj.canvas, hj.cls, hj.reg) + hj.counters())
    for j, hj in handles:
        got.append((j, hj.canvas, hj.cls, hj.reg) + hj.counters())
    assert len(got) == 6
    for j, c, cls, reg, n, k in got:
        assert torch.equal(c, want[j][0]) and torch.equal(cls, want[j][1]) and torch.equal(reg, want[j][2])
        assert torch.equal(n, want[j][3]) and torch.equal(k, want[j][4])


def _canvas_close(a, b, amp_scale):
    """PFN tolerance of test_gpu_pfn.py with a scalar bound on the conditioning magnitude."""
    a = a.double().cpu().numpy(); b = b.double().cpu().numpy()
    tol = 1e-5 * np.maximum(np.abs(a), np.abs(b)) + 2e-6 + 1e-6 * amp_scale
    err = np.abs(a - b) - tol
    assert err.max() <= 0, "max violation %g (max abs diff %g)" % (err.max(), np.abs(a - b).max())


@pytest.mark.parametrize("dense_mean", [True, False])
@pytest.mark.parametrize("training", [True, False])
def test_fused_input_path_equals_pillarize_then_encode(dense_mean, training):
    """pp_input_path (x never materialised, padding slots evaluated once per (p,n)) vs the
    signature-preserving pp_pillarize -> x -> pp_pfn_scatter sequence at the reference sizes:
    identical indices / counts, canvas within the PFN tolerance, identical running statistics up to
    that tolerance; and the optional x output is bit-identical to pillarize's."""
    from pp_b200 import pipeline, synth
    P, N = 24000, 200
    mean = synth.make_data_mean(P, N, dense=dense_mean)
    prm = synth.make_pfn_params(1, flip_gamma=True)
    mk = lambda: pipeline.InputPath(data_mean=mean, pfn_params=prm, training=training)
    pa, pb = mk(), mk()
    sweeps = [synth.make_sweep(40 + s) for s in range(3)]
    offs = np.cumsum([0] + [len(s) for s in sweeps]).tolist()
    pts = torch.from_numpy(np.concatenate(sweeps)).cuda()
    x, inds, npil = pa.pillarize(pts, offs)
    want = pa.encode(x, inds)
    canvas, inds2, npil2, x2 = pb.pillarize_encode(pts, offs, want_x=True)
    assert torch.equal(inds, inds2) and torch.equal(npil, npil2) and torch.equal(x, x2)
    # conditioning magnitude: |gamma|/sigma * max(sum_d |w x| + |b|) over the batch
    w = torch.from_numpy(prm["conv_w"]).cuda().abs()
    absdot = torch.einsum('cd,bdpn->bcpn', w, x[:, :, :2048].abs()).amax() + float(np.abs(prm["conv_b"]).max())
    var = pa.net.bn1.running_var if not training else None
    sigma = float(torch.sqrt(pa.net.bn1.running_var.min() + 1e-5)) if not training else 1.0
    amp = float(absdot) * float(np.abs(prm["bn_w"]).max()) / min(sigma, 1.0) * 4.0
    _canvas_close(canvas, want, amp)
    # support: identical up to elements whose value is an exact zero in one evaluation and below the absolute
    # tolerance in the other (BN(relu(.)) of a pillar extreme that cancels); each such element is checked
    diff = (canvas != 0) ^ (want != 0)
    n_diff = int(diff.sum())
    print("canvas support mismatches: %d of %d non-zeros" % (n_diff, int((want != 0).sum())))
    assert n_diff <= 10
    if n_diff:
        assert float(torch.maximum(canvas[diff].abs(), want[diff].abs()).max()) <= 2e-6 + 1e-6 * amp
    if training:
        assert torch.allclose(pa.net.bn1.running_mean, pb.net.bn1.running_mean, rtol=1e-5, atol=1e-6)
        assert torch.allclose(pa.net.bn1.running_var, pb.net.bn1.running_var, rtol=1e-5, atol=1e-6)
    from pp_b200 import _runtime
    _runtime.check_status(torch.device("cuda"), "fused input path")


def test_fused_input_path_against_fp64_oracle():
    """Small grid, many points per pillar (some over the N cap), hand-checkable sizes: fused path vs the
    fp64 oracle composition (oracle.glue + oracle.pfn)."""
    import pp_b200
    from oracle import glue, pfn
    from pp_b200 import pipeline, synth
    cfg = pp_b200.PPConfig(max_pillars=600, max_points_per_pillar=16)
    P, N = 600, 16
    mean = synth.make_data_mean(P, N, dense=True)
    prm = synth.make_pfn_params(5, flip_gamma=True)
    path = pipeline.InputPath(cfg, data_mean=mean, pfn_params=prm, training=True)
    rng = np.random.default_rng(3)
    sweeps = []
    for s in range(2):
        n = 6000
        pts = np.zeros((n, 5), np.float32)
        pts[:, 0] = rng.uniform(-6, 6, n); pts[:, 1] = rng.uniform(-6, 6, n)      # ~60x60 cells: > P pillars, some > N points
        pts[:, 2] = rng.uniform(-2, 2, n); pts[:, 3] = 100.0
        sweeps.append(pts)
    offs = np.cumsum([0] + [len(s) for s in sweeps]).tolist()
    canvas, inds, npil = path.pillarize_encode(torch.from_numpy(np.concatenate(sweeps)).cuda(), offs)
    xs, iis = [], []
    for s in sweeps:
        x1, i1 = glue.pillarize(s[:, :4].astype(np.float64), torch.from_numpy(mean), max_pillars=P, max_points=N)
        xs.append(x1); iis.append(i1)
    x = torch.stack(xs); ii = torch.stack(iis)
    t = lambda a: torch.from_numpy(a)
    y, rm, rv = pfn.pfn_forward(x, t(prm["conv_w"]), t(prm["conv_b"]), t(prm["bn_w"]), t(prm["bn_b"]),
                                t(prm["running_mean"]), t(prm["running_var"]), True)
    want = pfn.scatter(y.float(), ii, cfg.canvas_height, cfg.canvas_width)
    assert torch.equal(inds.cpu(), ii)
    absdot = torch.einsum('cd,bdpn->bcpn', t(prm["conv_w"]).abs().double(), x.abs().double()).amax()
    _canvas_close(canvas, want, float(absdot) * 4.0)
    assert torch.allclose(path.net.bn1.running_var.cpu(), rv.float(), rtol=1e-5, atol=1e-6)


def test_fused_input_path_full_size_against_fp64_oracle():
    """The reference shape (P = 24000, N = 200, C = 64, 600 x 600 canvas), one sweep, training-mode BatchNorm:
    pp_input_path against the float64 oracle (oracle.glue + the PPFeatureNet algebra of oracle.pfn, evaluated in
    pillar chunks so that the [64,P,N] float64 intermediate never exists: per-channel sums of relu(y) and
    relu(y)^2 and the per-(pillar, channel) maximum AND minimum of y in one pass, BatchNorm + the monotone-map
    selection afterwards -- the same values as pfn.pfn_forward, checked against it at the small size above)."""
    from oracle import glue
    from pp_b200 import pipeline, synth
    P, N = 24000, 200
    mean = synth.make_data_mean(P, N, dense=True)
    prm = synth.make_pfn_params(7, flip_gamma=True)
    path = pipeline.InputPath(data_mean=mean, pfn_params=prm, training=True)
    sweep = synth.make_sweep(31)
    canvas, inds, npil = path.pillarize_encode(torch.from_numpy(sweep).cuda(), [0, len(sweep)])
    x, ii = glue.pillarize(sweep[:, :4].astype(np.float64), torch.from_numpy(mean), max_pillars=P, max_points=N)
    assert torch.equal(inds[0].cpu(), ii)
    W = torch.from_numpy(prm["conv_w"]).double(); bias = torch.from_numpy(prm["conv_b"]).double()
    S = torch.zeros(64, dtype=torch.float64); Q = torch.zeros(64, dtype=torch.float64)
    ymax = torch.empty((64, P), dtype=torch.float64); ymin = torch.empty((64, P), dtype=torch.float64)
    amp_abs = 0.0
    step = 1500
    for p0 in range(0, P, step):
        xc = x[:, p0:p0 + step].double()                                    # [9, p, N]
        y = torch.einsum('cd,dpn->cpn', W, xc) + bias.view(-1, 1, 1)
        r = torch.relu(y)
        S += r.sum(dim=(1, 2)); Q += (r * r).sum(dim=(1, 2))
        ymax[:, p0:p0 + step] = y.amax(dim=2); ymin[:, p0:p0 + step] = y.amin(dim=2)
        amp_abs = max(amp_abs, float((torch.einsum('cd,dpn->cpn', W.abs(), xc.abs()).amax() + bias.abs().max())))
    M = float(P * N)
    mu = S / M
    var = Q / M - mu * mu
    g = torch.from_numpy(prm["bn_w"]).double(); beta = torch.from_numpy(prm["bn_b"]).double()
    scale = g / torch.sqrt(var + 1e-5)
    ext = torch.where((scale >= 0).view(-1, 1), ymax, ymin)               # max_n BN(relu(y)) by monotonicity
    out = (torch.relu(ext) - mu.view(-1, 1)) * scale.view(-1, 1) + beta.view(-1, 1)     # [64, P]
    want = torch.zeros((64, 600, 600), dtype=torch.float64)
    live = ii[:, 0] != 0
    want[:, ii[live, 2], ii[live, 1]] = out[:, live]
    amp = amp_abs * float((g.abs() / torch.sqrt(var + 1e-5)).max())
    got = canvas[0].double().cpu()
    tol = 1e-5 * torch.maximum(got.abs(), want.abs()) + 2e-6 + 1e-6 * amp
    viol = (got - want).abs() - tol
    plain = (got - want).abs() > 1e-5 * torch.maximum(got.abs(), want.abs())
    print("full size: %d non-zero outputs, %d outside plain 1e-5 relative, max |diff| %.3g, amp %.3g" % (
        int((want != 0).sum()), int(plain.sum()), float((got - want).abs().max()), amp))
    assert float(viol.max()) <= 0
    rv = 0.9 * torch.from_numpy(prm["running_var"]).double() + 0.1 * var * (M / (M - 1))
    rm = 0.9 * torch.from_numpy(prm["running_mean"]).double() + 0.1 * mu
    assert torch.allclose(path.net.bn1.running_var.double().cpu(), rv, rtol=1e-5, atol=1e-7)
    assert torch.allclose(path.net.bn1.running_mean.double().cpu(), rm, rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize("fused", [True, False])
def test_two_lane_training_steps_equal_sequential_steps(fused):
    """Training-mode BN updates the running statistics every step: the streaming lanes must give exactly
    the sequence of canvases and statistics that blocking steps give (encode stages are ordered)."""
    from pp_b200 import pipeline, synth
    mean = synth.make_data_mean(24000, 200, dense=True)
    mk = lambda: pipeline.InputPath(data_mean=mean, pfn_params=synth.make_pfn_params(2), training=True, fused=fused)
    pa, pb = mk(), mk()
    batches = []
    for k in range(4):
        sweeps = [synth.make_sweep(20 * k + s) for s in range(2)]
        gts = [synth.make_gt(20 * k + s, 15) for s in range(2)]
        batches.append(pa.pack_host_batch(sweeps, gts))
    want = []
    for b in batches:
        c = pa.step_host(b)[0]
        want.append((c.clone(), pa.net.bn1.running_var.clone()))
    handles = [pb.step_host_async(b) for b in batches]      # all four in flight
    for h, (c, rv) in zip(handles, want):
        h.synchronize()
        assert torch.equal(h.canvas, c)
    assert torch.equal(pb.net.bn1.running_var, want[-1][1])
    assert int(pb.net.bn1.num_batches_tracked) == 4


def test_fused_input_path_more_than_eight_sweeps():
    """B = 11 (two padding passes: 8 + 3 sweeps, statistics from the first only) vs the dense sequence."""
    import pp_b200
    from pp_b200 import pipeline, synth
    cfg = pp_b200.PPConfig(max_pillars=6000, max_points_per_pillar=64)
    P, N = 6000, 64
    mean = synth.make_data_mean(P, N, dense=True)
    prm = synth.make_pfn_params(6, flip_gamma=True)
    mk = lambda: pipeline.InputPath(cfg, data_mean=mean, pfn_params=prm, training=True)
    pa, pb = mk(), mk()
    sweeps = [synth.make_sweep(70 + s)[:30000] for s in range(11)]
    offs = np.cumsum([0] + [len(s) for s in sweeps]).tolist()
    pts = torch.from_numpy(np.concatenate(sweeps)).cuda()
    x, inds, npil = pa.pillarize(pts, offs)
    want = pa.encode(x, inds)
    canvas, inds2, npil2 = pb.pillarize_encode(pts, offs)
    assert torch.equal(inds, inds2) and torch.equal(npil, npil2)
    w = torch.from_numpy(prm["conv_w"]).cuda().abs()
    absdot = torch.einsum('cd,bdpn->bcpn', w, x[:, :, :1024].abs()).amax() + float(np.abs(prm["conv_b"]).max())
    _canvas_close(canvas, want, float(absdot) * float(np.abs(prm["bn_w"]).max()) * 4.0)
    assert torch.allclose(pa.net.bn1.running_var, pb.net.bn1.running_var, rtol=1e-5, atol=1e-6)
    d = (canvas.double() - want.double()).abs().max().item()
    assert d < 1e-3 * max(1.0, want.abs().max().item() * 1e-2), d


def test_fused_input_path_edge_cases():
    """Empty sweep, single-point sweep, everything out of range, one pillar holding more than N points:
    fused path == dense sequence (indices, counts, canvas within the PFN tolerance)."""
    import pp_b200
    from pp_b200 import pipeline, synth
    cfg = pp_b200.PPConfig(max_pillars=2000, max_points_per_pillar=24)
    P, N = 2000, 24
    mean = synth.make_data_mean(P, N, dense=True)
    prm = synth.make_pfn_params(8, flip_gamma=True)
    rng = np.random.default_rng(5)
    full = synth.make_sweep(90)[:20000]
    pile = np.zeros((100, 5), np.float32); pile[:, 0] = 3.03 + rng.uniform(0, .1, 100); pile[:, 1] = -7.01 - rng.uniform(0, .1, 100)
    pile[:, 2] = rng.uniform(-1, 1, 100); pile[:, 3] = 100.0                  # 100 points in one or two cells (> N)
    far = np.full((50, 5), 500.0, np.float32)                                 # all out of range
    one = np.array([[1.0, 2.0, 0.5, 100.0, 0.0]], np.float32)
    empty = np.zeros((0, 5), np.float32)
    sweeps = [full, empty, pile, far, one]
    offs = np.cumsum([0] + [len(s) for s in sweeps]).tolist()
    pts = torch.from_numpy(np.concatenate(sweeps)).cuda()
    for training in (True, False):
        mk = lambda: pipeline.InputPath(cfg, data_mean=mean, pfn_params=prm, training=training)
        pa, pb = mk(), mk()
        x, inds, npil = pa.pillarize(pts, offs)
        want = pa.encode(x, inds)
        canvas, inds2, npil2 = pb.pillarize_encode(pts, offs)
        assert torch.equal(inds, inds2) and torch.equal(npil, npil2)
        assert npil.tolist()[1] == 0 and npil.tolist()[3] == 0 and npil.tolist()[4] == 1
        assert not canvas[1].any() and not canvas[3].any() and not want[1].any()
        w = torch.from_numpy(prm["conv_w"]).cuda().abs()
        absdot = torch.einsum('cd,bdpn->bcpn', w, x.abs()).amax() + float(np.abs(prm["conv_b"]).max())
        _canvas_close(canvas, want, float(absdot) * float(np.abs(prm["bn_w"]).max()) * 4.0)
        assert torch.equal(canvas != 0, want != 0)
        if training:
            assert torch.allclose(pa.net.bn1.running_var, pb.net.bn1.running_var, rtol=1e-5, atol=1e-6)


def test_fused_dense_stress_cloud():
    """BASELINE config 5 shape through the fused path: 10-sweep cloud (~590k points), P = 30000 (the cap
    binds), 200 GT boxes; equals the dense sequence, targets included."""
    import pp_b200
    from pp_b200 import pipeline, synth
    cfg = pp_b200.PPConfig(max_pillars=30000)
    P, N = 30000, 200
    mean = synth.make_data_mean(P, N, dense=True)
    prm = synth.make_pfn_params(9)
    cloud = synth.make_dense_cloud(0) if hasattr(synth, "make_dense_cloud") else np.concatenate(
        [synth.make_sweep(100 + k) + np.array([0.3 * k, 0, 0, 0, 0], np.float32) for k in range(10)])
    gts = [synth.make_gt(3, 200)]
    pa = pipeline.InputPath(cfg, data_mean=mean, pfn_params=prm, training=True, fused=False)
    pb = pipeline.InputPath(cfg, data_mean=mean, pfn_params=prm, training=True, fused=True, anchors=pa.ensure_anchors())
    ba = pa.pack_host_batch([cloud], gts)
    ca, clsa, rega, na, ka = pa.step_host(ba)
    cb, clsb, regb, nb, kb = pb.step_host(pb.pack_host_batch([cloud], gts))
    assert int(na[0]) == 30000 and torch.equal(na, nb)
    assert torch.equal(clsa, clsb) and torch.equal(rega, regb) and torch.equal(ka, kb)
    d = (ca.double() - cb.double()).abs().max().item()
    assert d <= 1e-5 * ca.abs().max().item() + 2e-4, d
    assert torch.equal(ca != 0, cb != 0)
