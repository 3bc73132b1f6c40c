"""Generate tests/golden/targets_small.npz by running the REFERENCE's own target assignment.

Run in the build container only (needs /root/reference); the fixture travels, the reference does not.

    python tests/golden/make_golden_targets.py

What runs: ``/root/reference/utils/box_utils.py`` imported UNMODIFIED (oracle/refpy.py) -- its
``make_anchor_boxes`` (:111-159), ``boxes_to_image_space`` (:19-32), ``make_target`` (:70-109) and
``create_target`` (:162-232) -- with ``data.pillars.make_ious`` = the reference's ``data/pillars.cpp`` built
against the Boost stand-in (oracle/_ref) and the absent third-party packages replaced by the stand-ins of
oracle/sdk_shim (pyquaternion, lyft_dataset_sdk, easydict; their arithmetic is restated from memory of the
packages, see oracle/sdk_shim/README.md).  ``cfg.DATA.FM_HEIGHT/FM_WIDTH`` are set to 40 so that the dense
[A,G] path stays small (A = 9600); CANVAS_HEIGHT stays 600, as in tests/test_gpu_targets.py.

Cases (each asserted here, on the reference's own outputs, to actually hit its branch):
  rand_G12/G30/G60   seeded synthetic GT sets (pp_b200.synth.make_gt)
  anchor0            a GT whose best anchor is index 0: dropped by np.nonzero (:204); row 0 stays positive by threshold
  shared_top         two GTs of different class sharing one best anchor: two class bits, the later GT's reg row
  overwrite          an anchor positive for GT i (class c_i) that is also GT j's best anchor: row zeroed, only c_j set
  iou_eq_thresh      anchors whose max IoU is exactly 0.6 and that are nobody's best anchor: not positive (strict >)
  no_overlap         a GT far from every anchor: all-zero IoU column -> argmax 0 -> dropped, everything zero
  G == 0             the reference raises ValueError (np.max over an empty axis); recorded as a flag
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

FM = 40


def main():
    from oracle import refpy
    loaded = refpy.load()
    assert loaded is not None, "reference checkout (or oracle/_ref) not present"
    bu, cfg = loaded
    from pyquaternion import Quaternion                       # oracle/sdk_shim
    from lyft_dataset_sdk.utils.data_classes import Box       # oracle/sdk_shim
    import pp_b200
    from pp_b200 import synth

    cfg.DATA.FM_HEIGHT = FM
    cfg.DATA.FM_WIDTH = FM
    names = sorted(cfg.DATA.NAME_TO_IND, key=cfg.DATA.NAME_TO_IND.get)
    H = cfg.DATA.CANVAS_HEIGHT

    boxes, a_corners, a_centers, a_xy = bu.make_anchor_boxes()
    A = len(boxes)
    a_wlh = np.stack([b.wlh for b in boxes])
    a_yaw = np.array([b.orientation.yaw_pitch_roll[0] for b in boxes])
    out = {"fm": np.array([FM, FM]), "a_corners": a_corners, "a_centers": a_centers, "a_xy": a_xy,
           "a_wlh": a_wlh, "a_yaw": a_yaw, "class_names": np.array(names)}

    def gt_boxes(centers, wlh, yaw, cls):
        return [Box(list(c), list(s), Quaternion(axis=[0, 0, 1], radians=float(y)), name=names[int(k)])
                for c, s, y, k in zip(centers, wlh, yaw, cls)]

    def run(tag, centers, wlh, yaw, cls):
        """centers: canvas space, NOT flipped (what Box.center holds in the reference)."""
        centers, wlh = np.asarray(centers, np.float64), np.asarray(wlh, np.float64)
        g = gt_boxes(centers, wlh, yaw, cls)
        gc, gcor = bu.boxes_to_image_space(g)
        ious = np.zeros((A, len(g)))
        bu.pillars.make_ious(a_corners, gcor, a_centers, gc, ious)
        c, r = bu.create_target(a_corners, gcor, a_centers, gc, boxes, g)
        rows = np.nonzero((c != 0).any(1) | (r != 0).any(1))[0]
        nz = np.nonzero(ious)
        out.update({
            tag + "/g_centers": centers, tag + "/g_wlh": wlh, tag + "/g_cls": np.asarray(cls, np.int32),
            tag + "/g_yaw": np.array([b.orientation.yaw_pitch_roll[0] for b in g]),   # as the reference reads it
            tag + "/g_yaw_in": np.asarray(yaw, np.float64),
            tag + "/g_centers_img": gc, tag + "/g_corners_img": gcor,
            tag + "/rows": rows.astype(np.int64), tag + "/cls_rows": c[rows], tag + "/reg_rows": r[rows],
            tag + "/iou_a": nz[0].astype(np.int32), tag + "/iou_g": nz[1].astype(np.int32), tag + "/iou_v": ious[nz],
        })
        return c, r, ious, g

    # seeded synthetic sets, placed so that the flipped boxes land on the 80x80 corner of the canvas
    for seed, G in ((2, 12), (3, 30), (4, 60)):
        gt = synth.make_gt(seed, G, pp_b200.PPConfig(canvas_width=2 * FM, canvas_height=2 * FM))
        gt["centers"][:, 1] = (H - 1) - gt["centers"][:, 1]
        c, r, ious, _ = run("rand_G%d" % G, gt["centers"], gt["wlh"], gt["yaw"], gt["cls"])
        assert (ious.max(1) > 0.6).sum() > 0, "no positives in the seeded set"

    dims = [np.asarray(d, np.float64) for d in cfg.DATA.ANCHOR_DIMS]
    car, truck, ped, bike = (cfg.DATA.NAME_TO_IND[k] for k in ("car", "truck", "pedestrian", "bicycle"))
    aidx = lambda y, x, d: (y * FM + x) * 6 + d
    unflip = lambda c: [c[0], (H - 1) - c[1], c[2]]

    # --- anchor0: GT sitting exactly on anchor 0
    c, r, ious, _ = run("anchor0", [unflip(a_centers[0])], [dims[0]], [0.0], [bike])
    assert ious[:, 0].argmax() == 0 and ious[0, 0] > 0.6
    assert c[0, bike] == 1 and c.sum() == 1 and r[0, 0] == 1          # positive by threshold only
    # --- shared_top: identical boxes, different class
    ctr = [50.2, (H - 1) - 21.0, 0.5]
    c, r, ious, g = run("shared_top", [ctr, [ctr[0], ctr[1], 0.9]], [[9.0, 26.0, 1.7]] * 2, [0.05, 0.05], [car, truck])
    top = ious.argmax(0)
    assert top[0] == top[1] != 0 and c[top[0], car] == 1 and c[top[0], truck] == 1
    assert list(r[top[0]]) == bu.make_target(boxes[top[0]], g[1], top[0])
    # --- overwrite: seeded search for "positive for GT i, best anchor of GT j, different classes"
    found = None
    rng = np.random.default_rng(77)
    for trial in range(400):
        a = aidx(int(rng.integers(8, 30)), int(rng.integers(8, 30)), 2)
        ac = a_centers[a]
        gi = [ac[0] + rng.uniform(0.8, 1.6), ac[1] + rng.uniform(-0.3, 0.3), 0.6]        # image space
        gj = [ac[0] - rng.uniform(0.0, 0.6), ac[1] + rng.uniform(-0.2, 0.2), 0.4]
        wi = dims[2] * rng.uniform(0.97, 1.03, 3)
        wj = dims[2] * np.array([rng.uniform(0.45, 0.6), rng.uniform(0.75, 0.95), 1.0])
        cen, wl, yw, cl = [unflip(gi), unflip(gj)], [wi, wj], [0.01, -0.02], [car, truck]
        g = gt_boxes(np.asarray(cen), np.asarray(wl), yw, cl)
        gc, gcor = bu.boxes_to_image_space(g)
        io = np.zeros((A, 2))
        bu.pillars.make_ious(a_corners, gcor, a_centers, gc, io)
        top = io.argmax(0)
        aj = top[1]
        if aj != 0 and top[0] != aj and io[aj].argmax() == 0 and io[aj, 0] > 0.6:
            found = (cen, wl, yw, cl, aj)
            break
    assert found is not None, "no overwrite configuration found"
    cen, wl, yw, cl, aj = found
    c, r, ious, g = run("overwrite", cen, wl, yw, cl)
    assert ious[aj].argmax() == 0 and ious[aj, 0] > 0.6 and ious[:, 1].argmax() == aj
    assert c[aj, car] == 0 and c[aj, truck] == 1 and c[aj].sum() == 1
    assert list(r[aj]) == bu.make_target(boxes[aj], g[1], aj)
    # --- iou_eq_thresh: a 10 x 15 box inside every medium yaw-0 anchor within 5 units: IoU = 150/250 = 0.6 exactly
    y0 = 21
    c, r, ious, _ = run("iou_eq_thresh", [unflip([32.0, float(y0), 0.75])], [[10.0, 15.0, 1.75]], [0.0], [car])
    eq = np.nonzero(ious[:, 0] == 0.6)[0]
    assert len(eq) >= 4 and ious[:, 0].max() == 0.6 == cfg.DATA.IOU_POS_THRESH
    first = ious[:, 0].argmax()
    assert first == eq[0] and c[first, car] == 1                       # forced match on the first maximiser
    assert c[eq[1:]].sum() == 0 and r[eq[1:]].sum() == 0 and c.sum() == 1   # == threshold is NOT positive
    # --- no_overlap: beyond the centre prefilter of every anchor of the small lattice
    c, r, ious, _ = run("no_overlap", [unflip([300.0, 300.0, 0.0])], [dims[2]], [0.3], [car])
    assert not ious.any() and not c.any() and not r.any()
    # --- G == 0
    try:
        bu.create_target(a_corners, np.zeros((0, 4, 2)), a_centers, np.zeros((0, 3)), boxes, [])
        out["g0_raises"] = np.array(0)
    except ValueError:
        out["g0_raises"] = np.array(1)
    out["cases"] = np.array(["rand_G12", "rand_G30", "rand_G60", "anchor0", "shared_top", "overwrite",
                             "iou_eq_thresh", "no_overlap"])

    # --- make_target on seeded box pairs (every yaw branch: :92-95 and :99-102)
    rng = np.random.default_rng(5)
    n = 400
    mt_a = rng.integers(0, A, n)
    mt_c = np.stack([rng.uniform(0, 80, n), rng.uniform(H - 80, H, n), rng.uniform(-1, 2, n)], 1)
    mt_wlh = np.stack([d for d in dims])[rng.integers(0, 6, n)] * rng.uniform(0.5, 1.5, (n, 3))
    mt_yaw = rng.uniform(-np.pi, np.pi, n)
    mt_yaw[:8] = [np.pi / 2, -np.pi / 2, np.pi, -np.pi, 0.0, np.pi / 2 + 1e-9, np.pi / 2 - 1e-9, -np.pi / 2 + 1e-9]
    g = gt_boxes(mt_c, mt_wlh, mt_yaw, np.zeros(n, int))
    out["mt/anchor"] = mt_a.astype(np.int64)
    out["mt/g_centers"], out["mt/g_wlh"] = mt_c, mt_wlh
    out["mt/g_yaw"] = np.array([b.orientation.yaw_pitch_roll[0] for b in g])
    out["mt/target"] = np.array([bu.make_target(boxes[a], b, a) for a, b in zip(mt_a, g)], dtype=np.float64)
    assert set(out["mt/target"][:, 8]) == {0.0, 1.0}

    path = os.path.join(HERE, "targets_small.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", "G==0 raises:", int(out["g0_raises"]))


if __name__ == "__main__":
    main()
