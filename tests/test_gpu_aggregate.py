"""Device-side sweep aggregation (pp_aggregate_sweeps) vs oracle/aggregate.py: transformed
coordinates and the remove_close mask bit-exact; the pillars built from the device-aggregated cloud equal those
the oracle pipeline builds from the compacted cloud (dataset.py:54-106), bit for bit."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

SENTINEL = np.float32(3.0e38)


def _files(n_files, seed, n_pts=None):
    """Raw sensor-frame files plus poses like consecutive Lyft samples: the ego moves ~0.4 m and turns a little
    between sweeps; a handful of points sit within the remove_close radius after the transform."""
    from oracle import aggregate as og
    from pp_b200 import synth
    rng = np.random.default_rng(seed)
    sensor = og.pose_matrix([1.2, 0.0, 1.8], [0.999, 0.01, -0.02, 0.03])
    ref_pose = None
    raws, mats = [], []
    for k in range(n_files):
        pose = og.pose_matrix([100.0 - 0.4 * k, 50.0 + 0.05 * k, 0.2], [np.cos(0.2 - 0.004 * k), 0, 0, np.sin(0.2 - 0.004 * k)])
        if ref_pose is None:
            ref_pose = og.pose_matrix([100.0, 50.0, 0.2], [np.cos(0.2), 0, 0, np.sin(0.2)], inverse=True)
        M = ref_pose @ pose @ sensor                                            # dataset.py:78
        raw = synth.make_sweep(seed * 10 + k)
        if n_pts is not None:
            raw = raw[:n_pts]
        raw = raw.copy()
        # points that land within 1 mm of the reference car's origin: inverse-transform exact targets
        tgt = np.array([[0.0004, -0.0007, 1.0, 1], [0.0, 0.0, -2.0, 1], [0.00099, 0.00099, 0.3, 1], [0.0011, 0.0002, 0.5, 1]]).T
        raw[5:9, :3] = (np.linalg.inv(M) @ tgt)[:3].T.astype(np.float32)
        raws.append(raw); mats.append(M)
    return raws, mats


@pytest.mark.parametrize("n_files", [1, 3])
def test_kernel_matches_oracle_bit_for_bit(n_files):
    from oracle import aggregate as og
    from pp_b200 import pipeline
    raws, mats = _files(n_files, seed=4, n_pts=20000)
    want, keeps = og.aggregate(raws, mats)
    path = pipeline.InputPath(device=torch.device("cuda"))
    d = torch.tensor(np.concatenate(raws), device="cuda")
    offs = np.concatenate([[0], np.cumsum([len(r) for r in raws])])
    kept = path.aggregate(d, offs, np.stack(mats), want_kept=True)
    got = d.cpu().numpy()
    keep = got[:, 0] != SENTINEL
    assert np.array_equal(keep, np.concatenate(keeps))
    assert (~keep).sum() >= n_files * 2                                         # the planted close points were dropped
    assert np.array_equal(got[keep][:, :4].astype(np.float64), want)            # coordinates and intensity, exact
    assert np.array_equal(got[~keep][:, :3], np.full(((~keep).sum(), 3), SENTINEL))
    assert np.array_equal(got[:, 3:], np.concatenate(raws)[:, 3:])              # intensity / ring untouched
    assert kept.cpu().tolist() == [int(k.sum()) for k in keeps]


def test_pillars_of_the_device_aggregated_cloud_equal_the_oracle_pipeline():
    """Two samples in one batch: a 3-file aggregate and a single file, through pack_host_batch(transforms=...)
    -> upload (+ aggregate) -> pillarize, against oracle aggregate -> oracle create_pillars glue."""
    import pp_b200
    from oracle import aggregate as og, glue
    from pp_b200 import pipeline, synth
    cfg = pp_b200.PPConfig(max_pillars=6000, max_points_per_pillar=40)
    P, N = cfg.max_pillars, cfg.max_points_per_pillar
    mean = synth.make_data_mean(P, N, seed=1)
    path = pipeline.InputPath(cfg, data_mean=mean, device=torch.device("cuda"))
    ra, ma = _files(3, seed=7, n_pts=25000)
    rb, mb = _files(1, seed=8, n_pts=30000)
    batch = path.pack_host_batch([ra, rb], [synth.make_gt(1, 5), synth.make_gt(2, 5)], transforms=[ma, mb])
    assert batch["offsets"] == [0, 75000, 105000] and batch["n_files"] == 4
    d_pts, _ = path.upload(batch)
    x, inds, npil = path.pillarize(d_pts, batch["offsets"])
    torch.cuda.synchronize()
    for b, (raws, mats) in enumerate(((ra, ma), (rb, mb))):
        pts, _ = og.aggregate(raws, mats)
        xw, iw = glue.pillarize(pts, torch.from_numpy(mean), max_pillars=P, max_points=N)
        assert torch.equal(inds[b].cpu(), iw)
        assert torch.equal(x[b].cpu(), xw)
    assert int(npil.min()) > 1000


def test_ten_file_stress_cloud_through_the_fused_step():
    """BASELINE config 5 shape with the aggregation on the device: ten files (~590 k raw points) -> one sample,
    P = 30000; the canvas equals the one computed from the host-aggregated (oracle) cloud."""
    import pp_b200
    from oracle import aggregate as og
    from pp_b200 import pipeline, synth
    cfg = pp_b200.PPConfig(max_pillars=30000)
    mean = synth.make_data_mean(30000, 200, dense=True)
    prm = synth.make_pfn_params(9)
    raws, mats = _files(10, seed=2)
    gts = [synth.make_gt(3, 200)]
    pa = pipeline.InputPath(cfg, data_mean=mean, pfn_params=prm, training=True, fused=True)
    pb = pipeline.InputPath(cfg, data_mean=mean, pfn_params=prm, training=True, fused=True, anchors=pa.ensure_anchors())
    ca, clsa, rega, na, ka = pa.step_host(pa.pack_host_batch([raws], gts, transforms=[mats]))
    host_cloud, _ = og.aggregate(raws, mats)
    cb, clsb, regb, nb, kb = pb.step_host(pb.pack_host_batch([host_cloud.astype(np.float32)], gts))
    assert int(na[0]) == 30000 and torch.equal(na, nb)
    assert torch.equal(ca, cb) and torch.equal(clsa, clsb) and torch.equal(rega, regb)


def test_empty_file_in_the_middle_four_column_rows_and_no_points():
    from oracle import aggregate as og
    from pp_b200 import pipeline
    path = pipeline.InputPath(device=torch.device("cuda"))
    raws, mats = _files(3, seed=5, n_pts=3000)
    raws[1] = raws[1][:0]                                                       # a file without points
    want, keeps = og.aggregate(raws, mats)
    rows4 = np.ascontiguousarray(np.concatenate(raws)[:, :4])                  # S = 4: the float4 layout K1 prefers
    d = torch.tensor(rows4, device="cuda")
    kept = path.aggregate(d, [0, 3000, 3000, 6000], np.stack(mats), want_kept=True)
    got = d.cpu().numpy()
    keep = got[:, 0] != SENTINEL
    assert np.array_equal(keep, np.concatenate(keeps)) and kept.cpu().tolist() == [int(k.sum()) for k in keeps]
    assert np.array_equal(got[keep].astype(np.float64), want)
    empty = torch.empty((0, 5), dtype=torch.float32, device="cuda")
    assert path.aggregate(empty, [0, 0], np.eye(4)[None], want_kept=True).cpu().tolist() == [0]
