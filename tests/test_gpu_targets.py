"""K3 parity (through the C ABI): CUDA IoU / target assignment vs the oracle.

Labels (cls, reg column 0, ort) and per-GT best anchors are compared bit-exactly; IoU is expected
bit-identical to the oracle's IEEE-double evaluation (asserted <= 1e-12, bit-equality reported);
reg columns 1..7 within 1e-5 relative + 1e-7."""
import numpy as np
import pytest
import torch

from helpers import boxes_from_gt

pytestmark = pytest.mark.gpu


def _small_lattice(h=40, w=40):
    from oracle import targets as T
    boxes, corners, centers, _ = T.make_anchor_boxes(h, w)
    return boxes, corners, centers


def _gt_boxes(seed, G, size):
    import pp_b200
    from oracle import config as ocfg
    from oracle import targets as T
    from pp_b200 import synth
    gt = synth.make_gt(seed, G, pp_b200.PPConfig(canvas_width=size, canvas_height=size))
    # boxes are stored un-flipped; boxes_to_image_space flips with H=600 (config.py:60), so place them
    # such that the flipped boxes land on the small anchor lattice
    gt["centers"][:, 1] = 599 - gt["centers"][:, 1]
    return gt, boxes_from_gt(gt, T.Box, ocfg.CLASS_NAMES)


def test_make_ious_dropin_vs_oracle():
    from oracle import native, targets as T
    from pp_b200 import pillars
    boxes, corners, centers = _small_lattice()
    gt, g = _gt_boxes(1, 25, 80)
    gc, gcor = T.boxes_to_image_space(g)
    a = np.full((len(boxes), len(g)), -1.0); b = a.copy()
    native.make_ious(corners, gcor, centers, gc, a)
    pillars.make_ious(corners, gcor, centers, gc, b)
    assert np.abs(a - b).max() <= 1e-12
    print("make_ious bit-identical entries: %d / %d" % ((a == b).sum(), a.size))
    assert np.array_equal(a == 0, b == 0)
    assert (a > 0).sum() > 500


def test_wrong_winding_does_not_kill_the_process():
    """A GT ring given counter-clockwise: the reference may print 'IOU < 0' and exit(1)
    (data/pillars.cpp:166-169); here the value equals the oracle's and nothing exits."""
    from oracle import native
    from pp_b200 import pillars
    a = np.array([[[1, -1], [1, 1], [-1, 1], [-1, -1.0]]])
    ious = np.full((1, 1), -7.0); want = ious.copy()
    native.make_ious(a, a + 0.25, np.zeros((1, 3)), np.zeros((1, 3)), want)
    pillars.make_ious(a, a + 0.25, np.zeros((1, 3)), np.zeros((1, 3)), ious)       # GT given CCW
    assert ious[0, 0] == want[0, 0] <= 0.0


def _compare(cls, reg, c0, r0):
    assert np.array_equal(cls, c0.astype(np.float32))                       # labels bit-exact
    assert np.array_equal(reg[:, 0], r0[:, 0].astype(np.float32))
    assert np.array_equal(reg[:, 8], r0[:, 8].astype(np.float32))
    a, b = reg[:, 1:8].astype(np.float64), r0[:, 1:8].astype(np.float32).astype(np.float64)
    assert (np.abs(a - b) <= 1e-5 * np.maximum(np.abs(a), np.abs(b)) + 1e-7).all()
    return int((reg[:, 1:8] == r0[:, 1:8].astype(np.float32)).all(1).sum()), len(reg)


@pytest.mark.parametrize("seed,G", [(1, 1), (2, 12), (3, 30), (4, 60)])
def test_create_target_dropin_vs_oracle(seed, G):
    from oracle import targets as T
    from pp_b200 import box_utils
    boxes, corners, centers = _small_lattice()
    gt, g = _gt_boxes(seed, G, 80)
    gc, gcor = T.boxes_to_image_space(g)
    c0, r0, ious = T.create_target(corners, gcor, centers, gc, boxes, g, return_ious=True)
    cls, reg = box_utils.create_target(corners, gcor, centers, gc, boxes, g)
    assert cls.dtype == np.float64 and cls.shape == c0.shape and reg.shape == r0.shape
    _compare(cls.astype(np.float32), reg.astype(np.float32), c0, r0)
    near = int((np.abs(ious.max(1) - 0.6) < 1e-6).sum())
    print("G=%d positives=%d near-threshold anchors=%d" % (G, int((ious.max(1) > 0.6).sum()), near))


def test_branch_cases_anchor0_shared_best_and_threshold():
    from oracle import config as ocfg
    from oracle import targets as T
    from pp_b200 import box_utils
    boxes, corners, centers = _small_lattice()
    a0c = centers[0]
    g = [T.Box([31.3, 599 - 40.7, 0.0], [2.0, 2.5, 1.0], 0.4, "pedestrian"),
         T.Box([a0c[0], 599 - a0c[1], 0.5], ocfg.ANCHOR_DIMS[0], 0.0, "bicycle"),      # best anchor = index 0
         T.Box([50.2, 599 - 21.0, 0.5], [9.0, 26.0, 1.7], 0.05, "car"),
         T.Box([50.2, 599 - 21.0, 0.9], [9.0, 26.0, 1.7], 0.05, "truck"),               # same best anchor
         T.Box([500.0, 599 - 500.0, 0.0], [4.0, 9.0, 1.5], 2.5, "car")]                 # no anchor in reach
    gc, gcor = T.boxes_to_image_space(g)
    c0, r0 = T.create_target(corners, gcor, centers, gc, boxes, g)
    cls, reg = box_utils.create_target(corners, gcor, centers, gc, boxes, g)
    _compare(cls.astype(np.float32), reg.astype(np.float32), c0, r0)
    assert (c0.sum(1) == 2).any()


def test_batch_equals_per_sweep_and_empty_gt():
    import pp_b200
    from pp_b200 import box_utils
    cfg = pp_b200.PPConfig(fm_height=40, fm_width=40, canvas_height=80, canvas_width=80)
    anchors = box_utils.AnchorSet.from_config(cfg)
    from pp_b200 import synth
    gts = [synth.make_gt(s, G, cfg) for s, G in ((1, 7), (2, 0), (3, 19))]   # H=80 here: flips stay on the lattice
    packs = []
    for gt in gts:
        cen, cor = box_utils.gt_to_image_space(gt, cfg.canvas_height) if len(gt["yaw"]) else (np.zeros((0, 3)), np.zeros((0, 4, 2)))
        packs.append((cor, cen, gt["wlh"], gt["yaw"], gt["cls"]))
    f = lambda k, dt: torch.from_numpy(np.concatenate([np.asarray(p[k]).reshape((-1,) + np.asarray(p[k]).shape[1:]) for p in packs]).astype(dt)).cuda()
    offs = np.cumsum([0] + [len(p[3]) for p in packs]).tolist()
    cls, reg, top, counts = box_utils.assign_targets(anchors, f(0, np.float64), f(1, np.float64), f(2, np.float64),
                                                     f(3, np.float64), f(4, np.int32), offs, cfg.num_classes, cfg.iou_pos_thresh)
    assert not cls[1].any() and not reg[1].any()                 # sweep without GT: all zeros
    for b, p in enumerate(packs):
        if len(p[3]) == 0:
            continue
        g = lambda k, dt: torch.from_numpy(np.asarray(p[k]).astype(dt)).cuda()
        c1, r1, t1, n1 = box_utils.assign_targets(anchors, g(0, np.float64), g(1, np.float64), g(2, np.float64),
                                                  g(3, np.float64), g(4, np.int32), [0, len(p[3])], cfg.num_classes,
                                                  cfg.iou_pos_thresh)
        assert torch.equal(c1[0], cls[b]) and torch.equal(r1[0], reg[b])
        assert torch.equal(t1, top[offs[b]:offs[b + 1]]) and torch.equal(n1[0], counts[b])


def test_full_anchor_grid_config3():
    """BASELINE config 3: the full 540000-anchor lattice, 100 GT boxes."""
    import pp_b200
    from oracle import config as ocfg
    from oracle import targets as T
    from pp_b200 import box_utils, synth
    arr = box_utils.make_anchor_arrays()
    anchors = box_utils.AnchorSet(arr["corners"], arr["centers"], arr["wlh"], arr["yaw"])

    class _Lazy:                                   # anchor_box_list stand-in: only indexed for positives
        def __len__(self):
            return arr["centers"].shape[0]

        def __getitem__(self, a):
            return T.Box(arr["centers"][a], arr["wlh"][a], arr["yaw"][a])
    gt = synth.make_gt(7, 100)
    g = boxes_from_gt(gt, T.Box, ocfg.CLASS_NAMES)
    gc, gcor = T.boxes_to_image_space(g)
    c0, r0, ious = T.create_target(arr["corners"], gcor, arr["centers"], gc, _Lazy(), g, return_ious=True)
    cls, reg = box_utils.create_target(arr["corners"], gcor, arr["centers"], gc, anchors, g)
    same, n = _compare(cls.astype(np.float32), reg.astype(np.float32), c0, r0)
    npos = int((ious.max(1) > 0.6).sum())
    print("A=%d G=100 positives=%d forced=%d reg rows bit-identical=%d/%d near-threshold=%d" % (
        n, npos, int((ious.argmax(0) != 0).sum()), same, n, int((np.abs(ious.max(1) - 0.6) < 1e-6).sum())))
    assert npos > 20


# ---- against the REFERENCE's own utils/box_utils.py (tests/golden/targets_small.npz, made by
# ---- tests/golden/make_golden_targets.py: the reference module imported unmodified, run in the build container)
def _fixture_inputs(fx, tag):
    from helpers import fixture_case
    c = fixture_case(fx, tag)
    f = lambda a, dt=np.float64: torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).cuda()
    return c, (f(c["g_corners_img"]), f(c["g_centers_img"]), f(c["g_wlh"]), f(c["g_yaw"]), f(c["g_cls"], np.int32))


@pytest.mark.parametrize("tag", ["rand_G12", "rand_G30", "rand_G60", "anchor0", "shared_top", "overwrite",
                                 "iou_eq_thresh", "no_overlap"])
def test_assign_targets_equals_reference_fixture(tag):
    """pp_assign_targets and pp_assign_targets_list on the fixture's inputs: cls labels, reg columns 0 and 8
    bit for bit, reg columns 1..7 to 1e-5 relative + 1e-7 against what the reference's create_target returned."""
    from helpers import targets_fixture
    from pp_b200 import box_utils
    fx = targets_fixture()
    anchors = box_utils.AnchorSet(fx["a_corners"], fx["a_centers"], fx["a_wlh"], fx["a_yaw"])
    c, gt = _fixture_inputs(fx, tag)
    G = len(c["g_cls"])
    cls, reg, top, counts = box_utils.assign_targets(anchors, *gt, [0, G])
    torch.cuda.synchronize()
    same, n = _compare(cls[0].cpu().numpy(), reg[0].cpu().numpy(), c["cls"], c["reg"])
    # per-GT best anchors = the reference's np.argmax over the transposed matrix (first index on ties)
    np.testing.assert_array_equal(top.cpu().numpy(), c["ious"].argmax(0).astype(np.int32))
    pos, _, top2, _ = box_utils.assign_targets(anchors, *gt, [0, G], as_list=True)
    dc, dr = pos.dense()
    assert torch.equal(dc, cls) and torch.equal(dr, reg) and torch.equal(top2, top)
    if tag == "iou_eq_thresh":
        eq = np.nonzero(c["ious"][:, 0] == 0.6)[0]
        assert len(eq) >= 4 and not cls[0].cpu().numpy()[eq[1:]].any()      # == 0.6 is not positive
    print("%s: reg rows bit-identical %d/%d, flagged rows %d" % (tag, same, n, int((c["reg"][:, 0] != 0).sum())))


def test_create_target_dropin_equals_reference_fixture():
    """The numpy-signature drop-in (box_utils.create_target) called the way data/dataset.py:113-116 calls the
    reference: arrays + box lists."""
    from helpers import fixture_case, targets_fixture
    from oracle import targets as T
    from pp_b200 import box_utils
    fx = targets_fixture()
    names = [str(s) for s in fx["class_names"]]
    boxes = T.LazyAnchorBoxes(fx["a_centers"], fx["a_wlh"], fx["a_yaw"])
    boxes = [boxes[a] for a in range(len(boxes))]
    for tag in ("rand_G30", "overwrite"):
        c = fixture_case(fx, tag)
        g = [T.Box(c["g_centers"][i], c["g_wlh"][i], c["g_yaw"][i], names[int(c["g_cls"][i])]) for i in range(len(c["g_yaw"]))]
        cls, reg = box_utils.create_target(fx["a_corners"], c["g_corners_img"], fx["a_centers"], c["g_centers_img"], boxes, g)
        _compare(cls.astype(np.float32), reg.astype(np.float32), c["cls"], c["reg"])
