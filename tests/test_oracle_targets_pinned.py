"""The target-assignment oracle against the REFERENCE's own code.

tests/golden/targets_small.npz holds inputs and outputs of /root/reference/utils/box_utils.py run unmodified
(make_anchor_boxes, boxes_to_image_space, make_target, create_target; tests/golden/make_golden_targets.py,
oracle/refpy.py, third-party packages replaced by oracle/sdk_shim).  Here:
  * oracle/targets.py and oracle/pp_oracle.c reproduce every fixture output bit for bit;
  * where /root/reference is present, the reference is run live on fresh seeded cases and compared again;
  * the product-side anchor arrays / boxes_to_image_space equal the reference's (corners to 1 ulp: the SDK
    rotates through a quaternion rotation matrix, the product through cos/sin of the yaw).
"""
import numpy as np
import pytest

from helpers import fixture_case, targets_fixture
from oracle import config as ocfg
from oracle import native, refpy
from oracle import targets as T

FX = targets_fixture()
CASES = [str(c) for c in FX["cases"]]


def _oracle_boxes(centers, wlh, yaw, cls=None):
    names = [str(n) for n in FX["class_names"]]
    return [T.Box(centers[i], wlh[i], yaw[i], None if cls is None else names[int(cls[i])]) for i in range(len(yaw))]


def _anchor_boxes():
    return T.LazyAnchorBoxes(FX["a_centers"], FX["a_wlh"], FX["a_yaw"])


def test_class_table_matches_reference():
    assert [str(n) for n in FX["class_names"]] == list(ocfg.CLASS_NAMES)


@pytest.mark.parametrize("tag", CASES)
def test_oracle_create_target_equals_reference_fixture(tag):
    c = fixture_case(FX, tag)
    g = _oracle_boxes(c["g_centers"], c["g_wlh"], c["g_yaw"], c["g_cls"])
    cls, reg, ious = T.create_target(FX["a_corners"], c["g_corners_img"], FX["a_centers"], c["g_centers_img"],
                                     _anchor_boxes(), g, return_ious=True)
    np.testing.assert_array_equal(ious, c["ious"])        # pp_oracle.c make_ious == the reference's make_ious
    np.testing.assert_array_equal(cls, c["cls"])
    np.testing.assert_array_equal(reg, c["reg"])          # all nine columns, float64 bit for bit


def test_oracle_make_target_equals_reference_fixture():
    boxes = _anchor_boxes()
    g = _oracle_boxes(FX["mt/g_centers"], FX["mt/g_wlh"], FX["mt/g_yaw"])
    got = np.array([T.make_target(boxes[int(a)], b) for a, b in zip(FX["mt/anchor"], g)], dtype=np.float64)
    np.testing.assert_array_equal(got, FX["mt/target"])
    assert 30 < FX["mt/target"][:, 8].sum() < 370        # both orientation branches present


def test_branch_cases_hold_in_the_fixture():
    """The fixture really contains the branches it is named after (asserted on the reference's outputs)."""
    car, truck, bike = (ocfg.NAME_TO_IND[k] for k in ("car", "truck", "bicycle"))
    c = fixture_case(FX, "anchor0")
    assert c["ious"][:, 0].argmax() == 0 and c["cls"][0, bike] == 1 and c["cls"].sum() == 1
    c = fixture_case(FX, "shared_top")
    top = c["ious"].argmax(0)
    assert top[0] == top[1] and c["cls"][top[0], car] == 1 and c["cls"][top[0], truck] == 1
    c = fixture_case(FX, "overwrite")
    aj = c["ious"][:, 1].argmax()
    assert c["ious"][aj].argmax() == 0 and c["ious"][aj, 0] > 0.6 and c["cls"][aj, car] == 0 and c["cls"][aj, truck] == 1
    c = fixture_case(FX, "iou_eq_thresh")
    eq = np.nonzero(c["ious"][:, 0] == 0.6)[0]
    assert len(eq) >= 4 and c["cls"].sum() == 1 and c["cls"][eq[0], car] == 1 and not c["reg"][eq[1:]].any()
    c = fixture_case(FX, "no_overlap")
    assert not c["ious"].any() and not c["cls"].any()
    assert int(FX["g0_raises"]) == 1                       # the reference cannot take G == 0


def test_anchor_arrays_vs_reference_fixture():
    import pp_b200
    from pp_b200 import box_utils
    fm = FX["fm"]
    arr = box_utils.make_anchor_arrays(pp_b200.PPConfig(fm_height=int(fm[0]), fm_width=int(fm[1])))
    oc, octr, owlh, oyaw = T.anchor_arrays(int(fm[0]), int(fm[1]))
    for corners, centers, wlh, yaw in ((arr["corners"], arr["centers"], arr["wlh"], arr["yaw"]), (oc, octr, owlh, oyaw)):
        np.testing.assert_array_equal(centers, FX["a_centers"])
        np.testing.assert_array_equal(wlh, FX["a_wlh"])
        np.testing.assert_array_equal(yaw, FX["a_yaw"])   # incl. the 90-degree anchors' pi/2 - 2.2e-16
        # <= 1 ulp: the SDK's np.dot goes through the BLAS (FMA), not reproducible with plain arithmetic
        assert np.abs(corners - FX["a_corners"]).max() <= np.spacing(np.abs(corners).max())
        even = np.arange(len(yaw)) % 2 == 0               # yaw-0 anchors: identity rotation, exact
        np.testing.assert_array_equal(corners[even], FX["a_corners"][even])


@pytest.mark.parametrize("tag", ["rand_G30", "overwrite"])
def test_boxes_to_image_space_vs_reference_fixture(tag):
    from pp_b200 import box_utils
    c = fixture_case(FX, tag)
    g = _oracle_boxes(c["g_centers"], c["g_wlh"], c["g_yaw"], c["g_cls"])
    for fn in (T.boxes_to_image_space, box_utils.boxes_to_image_space):
        gc, gcor = fn(g)
        np.testing.assert_array_equal(gc, c["g_centers_img"])
        assert np.abs(gcor - c["g_corners_img"]).max() <= 2e-13    # cos/sin of the yaw vs the quaternion matrix
    gc, gcor = box_utils.gt_to_image_space({"centers": c["g_centers"], "wlh": c["g_wlh"], "yaw": c["g_yaw"]})
    np.testing.assert_array_equal(gc, c["g_centers_img"])
    assert np.abs(gcor - c["g_corners_img"]).max() <= 2e-13


@pytest.mark.skipif(not refpy.available(), reason="reference checkout not present (GPU box)")
@pytest.mark.parametrize("seed,G", [(11, 5), (12, 40), (13, 100)])
def test_reference_run_live_equals_oracle(seed, G):
    """Fresh seeded GT sets through the reference's create_target (run here) and through the restatement."""
    import pp_b200
    from pp_b200 import synth
    loaded = refpy.load()
    if loaded is None:
        pytest.skip("oracle/_ref not built")
    bu, cfg = loaded
    from pyquaternion import Quaternion
    from lyft_dataset_sdk.utils.data_classes import Box
    fm = int(FX["fm"][0])
    cfg.DATA.FM_HEIGHT = cfg.DATA.FM_WIDTH = fm
    if not hasattr(test_reference_run_live_equals_oracle, "_anchors"):
        test_reference_run_live_equals_oracle._anchors = bu.make_anchor_boxes()
    boxes, a_corners, a_centers, _ = test_reference_run_live_equals_oracle._anchors
    np.testing.assert_array_equal(a_corners, FX["a_corners"])
    gt = synth.make_gt(seed, G, pp_b200.PPConfig(canvas_width=2 * fm, canvas_height=2 * fm))
    gt["centers"][:, 1] = 599 - gt["centers"][:, 1]
    names = list(ocfg.CLASS_NAMES)
    g_ref = [Box(list(gt["centers"][i]), list(gt["wlh"][i]), Quaternion(axis=[0, 0, 1], radians=float(gt["yaw"][i])),
                 name=names[int(gt["cls"][i])]) for i in range(G)]
    gc, gcor = bu.boxes_to_image_space(g_ref)
    c_ref, r_ref = bu.create_target(a_corners, gcor, a_centers, gc, boxes, g_ref)
    yaw_read = np.array([b.orientation.yaw_pitch_roll[0] for b in g_ref])
    g_or = _oracle_boxes(gt["centers"], gt["wlh"], yaw_read, gt["cls"])
    c_or, r_or = T.create_target(a_corners, gcor, a_centers, gc, _anchor_boxes(), g_or)
    np.testing.assert_array_equal(c_or, c_ref)
    np.testing.assert_array_equal(r_or, r_ref)
    assert c_ref.sum() >= 1
