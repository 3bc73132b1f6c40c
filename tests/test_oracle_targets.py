"""Target-assignment oracle (numpy restatement of utils/box_utils.py): branch cases of
create_target / make_target, and the product-side anchor arrays against the loop version."""
import numpy as np
import pytest

from oracle import config as ocfg
from oracle import targets as T
from helpers import boxes_from_gt


@pytest.fixture(scope="module")
def small_anchors():
    # 40x40 feature map keeps the numpy [A,G] path quick: A = 9600
    boxes, corners, centers, xy = T.make_anchor_boxes(40, 40)
    return boxes, corners, centers


def _gt(centers, wlh, yaw, names):
    return [T.Box(c, s, y, n) for c, s, y, n in zip(centers, wlh, yaw, names)]


def test_anchor_lattice_layout(small_anchors):
    boxes, corners, centers = small_anchors
    assert corners.shape == (9600, 4, 2) and centers.shape == (9600, 3)
    a = (3 * 40 + 5) * 6 + 2          # y=3, x=5, d=2 (medium, yaw 0)
    np.testing.assert_array_equal(centers[a], [11.0, 7.0, 0.75])
    np.testing.assert_array_equal(boxes[a].wlh, ocfg.ANCHOR_DIMS[2])
    # counter-clockwise ring: positive shoelace area = w*l
    x, y = corners[a, :, 0], corners[a, :, 1]
    area = 0.5 * np.sum(x * np.roll(y, -1) - np.roll(x, -1) * y)
    assert area == pytest.approx(10.0 * 25.0)
    # yaw 90: length along y
    a90 = a + 1
    assert np.ptp(corners[a90, :, 1]) == pytest.approx(25.0)


def test_product_anchor_arrays_equal_loop_version():
    import pp_b200
    from pp_b200 import box_utils
    c = pp_b200.PPConfig(fm_height=12, fm_width=9)
    arr = box_utils.make_anchor_arrays(c)
    boxes, corners, centers, _ = T.make_anchor_boxes(12, 9)
    np.testing.assert_array_equal(arr["corners"], corners)
    np.testing.assert_array_equal(arr["centers"], centers)
    np.testing.assert_array_equal(arr["wlh"], np.stack([b.wlh for b in boxes]))
    np.testing.assert_array_equal(arr["yaw"], np.array([b.yaw for b in boxes]))


def test_make_target_formula():
    a = T.Box([11.0, 7.0, 0.75], [10.0, 25.0, 1.75], 0.0)
    g = T.Box([12.0, 590.0, 1.0], [9.0, 24.0, 1.5], 2.0, "car")     # yaw in [pi/2, pi] -> wrapped by -pi
    t = T.make_target(a, g)
    ad = np.sqrt(10.0 ** 2 + 25.0 ** 2)
    assert t[0] == 1
    assert t[1] == (12.0 - 11.0) / ad
    assert t[2] == ((599 - 590.0) - 7.0) / ad                       # flipped GT y, :83
    assert t[3] == (1.0 - 0.75) / 1.75
    assert t[4] == np.log(9.0 / 10.0) and t[5] == np.log(24.0 / 25.0) and t[6] == np.log(1.5 / 1.75)
    assert t[7] == np.sin((2.0 - np.pi) - 0.0)
    assert t[8] == 0
    a90 = T.Box([11.0, 7.0, 0.75], [10.0, 25.0, 1.75], np.pi / 2)
    assert T.make_target(a90, T.Box([0, 0, 0], [1, 1, 1], -0.2, "car"))[8] == 1   # dθ in [-pi,-pi/2]


def test_positive_threshold_is_strict_and_no_gt(small_anchors):
    boxes, corners, centers = small_anchors
    cls, reg = T.create_target(corners, np.zeros((0, 4, 2)), centers, np.zeros((0, 3)), boxes, []) \
        if False else (None, None)   # the reference cannot take G == 0 (np.max over an empty axis raises)
    a = (10 * 40 + 10) * 6 + 2
    # identical box -> IoU 1 -> positive; class bit of the GT
    g = _gt([[centers[a, 0], 599 - centers[a, 1], 0.75]], [ocfg.ANCHOR_DIMS[2]], [0.0], ["car"])
    gc, gcor = T.boxes_to_image_space(g)
    cls, reg, ious = T.create_target(corners, gcor, centers, gc, boxes, g, return_ious=True)
    assert ious[a, 0] == pytest.approx(1.0, abs=1e-14)
    assert cls[a, ocfg.NAME_TO_IND["car"]] == 1 and cls[a].sum() == 1
    assert reg[a, 0] == 1 and np.allclose(reg[a, 1:8], 0, atol=1e-15)
    pos = np.where(ious.max(1) > 0.6)[0]
    assert set(np.nonzero(reg[:, 0])[0]) == set(pos) | {a}


def test_forced_match_overrides_and_anchor0_dropped(small_anchors):
    boxes, corners, centers = small_anchors
    # GT 0: tiny box far from every anchor's threshold -> no positive, forced match on its best anchor
    # GT 1: sits on anchor 0 (y=0,x=0,d=0) -> its best anchor is index 0 -> dropped (np.nonzero, :204)
    # GT 2 and 3: identical boxes of different class -> share one best anchor -> two class bits, later reg wins
    a0c = centers[0]
    g = _gt([[31.3, 599 - 40.7, 0.0], [a0c[0], 599 - a0c[1], 0.5], [50.2, 599 - 21.0, 0.5], [50.2, 599 - 21.0, 0.9]],
            [[2.0, 2.5, 1.0], ocfg.ANCHOR_DIMS[0], [9.0, 26.0, 1.7], [9.0, 26.0, 1.7]],
            [0.4, 0.0, 0.05, 0.05], ["pedestrian", "bicycle", "car", "truck"])
    gc, gcor = T.boxes_to_image_space(g)
    cls, reg, ious = T.create_target(corners, gcor, centers, gc, boxes, g, return_ious=True)
    top = ious.argmax(0)
    assert ious[:, 0].max() < 0.6 and top[0] != 0
    assert cls[top[0], ocfg.NAME_TO_IND["pedestrian"]] == 1 and reg[top[0], 0] == 1
    assert top[1] == 0                                         # dropped: row 0 untouched by the forced pass
    assert ious[0, 1] > 0.6 and cls[0, ocfg.NAME_TO_IND["bicycle"]] == 1   # ... but positive by threshold
    assert top[2] == top[3]
    assert cls[top[2], ocfg.NAME_TO_IND["car"]] == 1 and cls[top[2], ocfg.NAME_TO_IND["truck"]] == 1
    np.testing.assert_array_equal(reg[top[2]], T.make_target(boxes[top[2]], g[3]))


def test_create_target_with_reference_build_ious(small_anchors, ref_module):
    if ref_module is None:
        pytest.skip("oracle/_ref not built")
    import pp_b200
    from pp_b200 import synth
    boxes, corners, centers = small_anchors
    gt = synth.make_gt(3, 12, pp_b200.PPConfig(canvas_width=80, canvas_height=80))
    g = boxes_from_gt(gt, T.Box, ocfg.CLASS_NAMES)
    gc, gcor = T.boxes_to_image_space(g)
    c0, r0 = T.create_target(corners, gcor, centers, gc, boxes, g)
    c1, r1 = T.create_target(corners, gcor, centers, gc, boxes, g, make_ious=ref_module.make_ious)
    np.testing.assert_array_equal(c0, c1)
    np.testing.assert_array_equal(r0, r1)
