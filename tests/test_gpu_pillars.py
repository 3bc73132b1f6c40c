"""K1 parity (through the C ABI): CUDA pillarize vs the CPU oracle.  Bit-exact for pillar
indices, membership, order, every decorated feature (fp64 drop-in) and the fp32 network tensor."""
import numpy as np
import pytest
import torch

from helpers import (GRID, cloud_boundaries, cloud_dense_cells, cloud_random, f32_exact,
                     run_create_pillars)

pytestmark = pytest.mark.gpu


def _both(pts, P, N, grid=GRID):
    from oracle import native
    from pp_b200 import pillars
    t0, i0 = run_create_pillars(native.create_pillars, pts, P, N, grid)
    t1, i1 = run_create_pillars(pillars.create_pillars, pts, P, N, grid)
    return (t0, i0), (t1, i1)


@pytest.mark.parametrize("case", ["boundaries", "random", "dense", "strided", "capP", "empty", "allout"])
def test_create_pillars_dropin_bit_exact(case):
    P, N = 3000, 20
    if case == "boundaries":
        pts, P, N = cloud_boundaries(), 16, 4
    elif case == "random":
        pts = cloud_random(11, n=6000)
    elif case == "dense":
        pts = cloud_dense_cells(12)
    elif case == "strided":
        pts = np.ascontiguousarray(cloud_random(13, n=3000, cols=4).T).T
    elif case == "capP":
        pts, P = cloud_random(14, n=6000, spread=59.0), 500
    elif case == "empty":
        pts = np.zeros((0, 4))
    else:
        pts = f32_exact([[100, 0, 0, 1], [0, -100, 0, 1], [0, 0, 11, 1]])
    (t0, i0), (t1, i1) = _both(pts, P, N)
    np.testing.assert_array_equal(i1, i0)
    np.testing.assert_array_equal(t1, t0)


def test_only_touched_slots_are_written():
    """In-place contract of data/pillars.cpp:48-56,390-392: untouched slots keep the caller's data."""
    from oracle import native
    from pp_b200 import pillars
    pts = cloud_random(15, n=500)
    P, N = 64, 3
    t0 = np.full((P, N, 9), 7.0); i0 = np.full((P, 3), 5.0)
    t1 = t0.copy(); i1 = i0.copy()
    native.create_pillars(pts, t0, i0, N, P, *GRID)
    pillars.create_pillars(pts, t1, i1, N, P, *GRID)
    np.testing.assert_array_equal(t1, t0)
    np.testing.assert_array_equal(i1, i0)
    assert (t1 == 7.0).any()


def test_pillar_with_more_points_than_the_block_scan_threshold():
    """> 1024 points in one cell takes the k_rank_big path; order and running mean stay exact."""
    rng = np.random.default_rng(5)
    n = 5000
    pts = np.stack([rng.uniform(3.01, 3.19, n), rng.uniform(-7.19, -7.01, n), rng.uniform(-2, 2, n),
                    rng.uniform(0, 1, n)], 1)
    pts[::7, 0] += 0.2                      # a second, interleaved pillar
    pts = f32_exact(pts)
    (t0, i0), (t1, i1) = _both(pts, 8, 200)
    assert int(i0[:, 0].sum()) == 2
    np.testing.assert_array_equal(i1, i0)
    np.testing.assert_array_equal(t1, t0)


def test_pillar_sizes_at_the_kernel_path_boundaries():
    """Pillars of exactly 1, 2, 32, 33 (thread -> warp / group paths of the rank and running-mean stages), 128, 129,
    256, 257 (group size of the shared-memory ranking), 1024 and 1025 points (block-scan path), interleaved in the
    input so that every segment is filled out of order: membership, order, running mean and features stay exact."""
    rng = np.random.default_rng(21)
    sizes = [1, 2, 32, 33, 128, 129, 256, 257, 1024, 1025]
    chunks = []
    for k, c in enumerate(sizes):
        x0, y0 = -50.0 + 3.0 * k, 10.0 + 2.0 * (k % 3)            # one 0.2 x 0.2 cell each
        chunks.append(np.stack([rng.uniform(x0 + 0.01, x0 + 0.19, c), rng.uniform(y0 + 0.01, y0 + 0.19, c),
                                rng.uniform(-2, 2, c), rng.uniform(0, 1, c)], 1))
    pts = np.concatenate(chunks)
    pts = f32_exact(pts[rng.permutation(len(pts))])
    for N in (200, 40):
        (t0, i0), (t1, i1) = _both(pts, 16, N)
        assert int(i0[:, 0].sum()) == len(sizes)
        np.testing.assert_array_equal(i1, i0)
        np.testing.assert_array_equal(t1, t0)


def test_non_float32_representable_doubles():
    """The drop-in takes arbitrary doubles, like the reference (binning in fp64)."""
    rng = np.random.default_rng(6)
    pts = rng.uniform(-61, 61, (4000, 4))
    pts[:50, 0] = np.nextafter(60.0, -np.inf)         # just below x_max
    pts[50:100, 1] = -60.0 + np.arange(50) * 0.2      # exactly on cell edges
    (t0, i0), (t1, i1) = _both(pts, 4000, 5)
    np.testing.assert_array_equal(i1, i0)
    np.testing.assert_array_equal(t1, t0)


def test_other_grid():
    grid = (.25, .5, -20, -40, -3, 30, 40, 1, 160)
    pts = cloud_random(16, n=5000, spread=45.0)
    (t0, i0), (t1, i1) = _both(pts, 2000, 7, grid)
    np.testing.assert_array_equal(i1, i0)
    np.testing.assert_array_equal(t1, t0)


@pytest.mark.parametrize("with_mean", [False, True])
def test_dense_batch_path_equals_dataset_glue(with_mean):
    """pp_pillarize (fp32 [B,9,P,N] with fused '- data_mean') vs data/dataset.py:88-106 on the oracle."""
    import pp_b200
    from oracle import glue
    from pp_b200 import pipeline, synth
    cfg = pp_b200.PPConfig(max_pillars=1500, max_points_per_pillar=24)
    P, N = cfg.max_pillars, cfg.max_points_per_pillar
    mean = synth.make_data_mean(P, N, seed=1) if with_mean else None
    path = pipeline.InputPath(cfg, data_mean=mean)
    sweeps = [synth.make_sweep(3)[:9000], synth.make_sweep(4)[:7000], np.zeros((0, 5), np.float32),
              synth.make_sweep(5)[:12000]]
    offs = np.cumsum([0] + [len(s) for s in sweeps]).tolist()
    d_pts = torch.from_numpy(np.concatenate(sweeps)).cuda()
    x, inds, npil = path.pillarize(d_pts, offs)
    torch.cuda.synchronize()
    for b, s in enumerate(sweeps):
        want_x, want_i = glue.pillarize(s.astype(np.float64), None if mean is None else torch.from_numpy(mean),
                                        P, N)
        assert int(npil[b]) == int(want_i[:, 0].sum())
        assert torch.equal(inds[b].cpu(), want_i)
        assert torch.equal(x[b].cpu(), want_x), "sweep %d" % b     # bit-exact fp32
    assert int(npil[1]) == P or int(npil[3]) == P                  # the P cap binds for one sweep


def test_full_size_sweep_bit_exact_and_properties():
    """BASELINE config sizes: P=24000, N=200, ~67k points (and float64 input through the same path)."""
    import pp_b200
    from oracle import glue
    from pp_b200 import pipeline, synth
    cfg = pp_b200.PPConfig()
    path = pipeline.InputPath(cfg)
    s = synth.make_sweep(1)
    x, inds, npil = path.pillarize(torch.from_numpy(s).cuda(), [0, len(s)])
    want_x, want_i = glue.pillarize(s.astype(np.float64))
    assert torch.equal(inds[0].cpu(), want_i)
    assert torch.equal(x[0].cpu(), want_x)
    # size-independent properties
    n = int(npil[0]); ii = inds[0].cpu().numpy()
    cells = ii[:n, 1] * 1000 + ii[:n, 2]
    assert len(np.unique(cells)) == n and np.all(ii[n:] == 0)
    xs = x[0].cpu().numpy()
    occ = np.any(xs != 0, axis=0)                              # [P,N] occupied slots
    assert not occ[n:].any()
    cnt = occ.sum(1)
    assert np.all(occ == (np.arange(cfg.max_points_per_pillar)[None, :] < cnt[:, None]))   # front-packed
    inr = (np.abs(s[:, 0]) < 60) & (np.abs(s[:, 1]) < 60) & (s[:, 2] >= -10) & (s[:, 2] < 10)
    assert cnt.sum() <= inr.sum()
    x64, i64, n64 = path.pillarize(torch.from_numpy(s[:, :4].astype(np.float64)).cuda(), [0, len(s)])
    assert torch.equal(x64, x) and torch.equal(i64, inds)


def test_dense_stress_cloud_cap_binds():
    """BASELINE config 5 shape: 10-sweep cloud (~590k points), P=30000: the first-touch P cap binds."""
    import pp_b200
    from oracle import glue
    from pp_b200 import pipeline, synth
    cfg = pp_b200.PPConfig(max_pillars=30000)
    path = pipeline.InputPath(cfg)
    s = synth.make_sweep(100, n_sweeps=10)
    x, inds, npil = path.pillarize(torch.from_numpy(s).cuda(), [0, len(s)])
    assert int(npil[0]) == 30000
    want_x, want_i = glue.pillarize(s.astype(np.float64), max_pillars=30000)
    assert torch.equal(inds[0].cpu(), want_i)
    assert torch.equal(x[0].cpu(), want_x)


def test_make_means_equals_reference_recurrence():
    """pp_b200.make_means (the equivalent of the reference's make_means.py:28-37 over this pillarizer) against the
    same recurrence over the oracle's network input."""
    import pp_b200
    from oracle import glue
    from pp_b200 import make_means, synth
    cfg = pp_b200.PPConfig(max_pillars=800, max_points_per_pillar=16)
    batches = [[synth.make_sweep(10 * b + s)[:5000] for s in range(3)] for b in range(3)]
    got = make_means.make_means(batches, cfg)
    means = torch.zeros(9 * 800 * 16)
    for i, sweeps in enumerate(batches):
        p = torch.stack([glue.pillarize(s[:, :4].astype(np.float64), None, max_pillars=800, max_points=16)[0] for s in sweeps])
        m = torch.mean(p.reshape(p.shape[0], -1), dim=0)
        means = means * (i / (i + 1)) + m * (1 / (i + 1))
    assert got.shape == means.shape and got.dtype == torch.float32
    assert torch.allclose(got, means, rtol=1e-6, atol=1e-6)      # torch.mean on the GPU vs CPU: summation order only
    assert float(got.abs().max()) > 1.0
