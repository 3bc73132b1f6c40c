"""Oracle (and product-side config) against the committed fixtures that were generated from the
reference itself (tests/golden/make_golden.py)."""
import json
import os

import numpy as np
import torch

from conftest import GOLDEN
from helpers import GRID, run_create_pillars
from oracle import config as ocfg
from oracle import native, pfn


def test_config_values_match_reference():
    g = json.load(open(os.path.join(GOLDEN, "config.json")))
    assert (ocfg.X_MIN, ocfg.Y_MIN, ocfg.Z_MIN, ocfg.X_MAX, ocfg.Y_MAX, ocfg.Z_MAX) == (
        g["X_MIN"], g["Y_MIN"], g["Z_MIN"], g["X_MAX"], g["Y_MAX"], g["Z_MAX"])
    assert (ocfg.X_STEP, ocfg.Y_STEP, ocfg.FM_SCALE) == (g["X_STEP"], g["Y_STEP"], g["FM_SCALE"])
    assert (int(ocfg.FM_HEIGHT), int(ocfg.FM_WIDTH), int(ocfg.CANVAS_HEIGHT), int(ocfg.CANVAS_WIDTH)) == (
        g["FM_HEIGHT"], g["FM_WIDTH"], g["CANVAS_HEIGHT"], g["CANVAS_WIDTH"])
    np.testing.assert_array_equal(np.stack(ocfg.ANCHOR_DIMS), np.array(g["ANCHOR_DIMS"]))
    assert list(ocfg.ANCHOR_YAWS) == g["ANCHOR_YAWS"] and list(ocfg.ANCHOR_ZS) == g["ANCHOR_ZS"]
    assert ocfg.MAX_PILLARS == g["MAX_PILLARS"] and ocfg.MAX_POINTS_PER_PILLAR == g["MAX_POINTS_PER_PILLAR"]
    assert ocfg.IOU_POS_THRESH == g["IOU_POS_THRESH"] and ocfg.NAME_TO_IND == g["NAME_TO_IND"]
    # the product-side defaults must be the same numbers
    import pp_b200
    c = pp_b200.PPConfig()
    assert (c.x_min, c.y_min, c.z_min, c.x_max, c.y_max, c.z_max) == (
        g["X_MIN"], g["Y_MIN"], g["Z_MIN"], g["X_MAX"], g["Y_MAX"], g["Z_MAX"])
    assert (c.x_step, c.y_step, c.fm_scale) == (g["X_STEP"], g["Y_STEP"], g["FM_SCALE"])
    assert (c.fm_height, c.fm_width, c.canvas_height, c.canvas_width) == (
        g["FM_HEIGHT"], g["FM_WIDTH"], g["CANVAS_HEIGHT"], g["CANVAS_WIDTH"])
    np.testing.assert_array_equal(np.stack(c.anchor_dims), np.array(g["ANCHOR_DIMS"]))
    assert list(c.anchor_yaws_deg) == g["ANCHOR_YAWS"] and list(c.anchor_zs) == g["ANCHOR_ZS"]
    assert (c.max_pillars, c.max_points_per_pillar, c.iou_pos_thresh, c.num_classes) == (
        g["MAX_PILLARS"], g["MAX_POINTS_PER_PILLAR"], g["IOU_POS_THRESH"], g["NUM_CLASSES"])
    assert c.name_to_ind == g["NAME_TO_IND"]


def test_create_pillars_and_ious_match_reference_fixture():
    g = np.load(os.path.join(GOLDEN, "pillars_small.npz"))
    for name in ("boundaries", "random", "dense", "capP"):
        pts = g[name + "/points"]
        P, N = g[name + "/PN"]
        t, ind = run_create_pillars(native.create_pillars, pts, int(P), int(N))
        np.testing.assert_array_equal(ind, g[name + "/indices"])
        want = np.zeros_like(t)
        idx = g[name + "/tensor_nz_index"]
        want[idx[:, 0], idx[:, 1], idx[:, 2]] = g[name + "/tensor_nz_value"]
        np.testing.assert_array_equal(t, want)
    ious = np.zeros_like(g["iou/ious"])
    native.make_ious(g["iou/a_corners"], g["iou/g_corners"], g["iou/a_centers"], g["iou/g_centers"], ious)
    np.testing.assert_array_equal(ious, g["iou/ious"])


def _run_pfn_oracle(g, tag, training):
    sd = {k.split("/", 2)[2]: torch.from_numpy(g[k]) for k in g.files if k.startswith(tag + "/sd0/")}
    x = torch.from_numpy(g["x"])
    return pfn.pfn_forward(x, sd["conv1.weight"], sd["conv1.bias"], sd["bn1.weight"], sd["bn1.bias"],
                           sd["bn1.running_mean"], sd["bn1.running_var"], training)


def test_pfn_oracle_matches_reference_module():
    g = np.load(os.path.join(GOLDEN, "pfn_small.npz"))
    for tag in ("pos", "mixed"):
        y, _, _ = _run_pfn_oracle(g, tag, False)
        np.testing.assert_allclose(y.numpy(), g[tag + "/y_eval"], rtol=1e-12, atol=1e-12)
        y, rm, rv = _run_pfn_oracle(g, tag, True)
        np.testing.assert_allclose(y.numpy(), g[tag + "/y_train"], rtol=1e-11, atol=1e-11)
        np.testing.assert_allclose(rm.numpy(), g[tag + "/rm1"], rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(rv.numpy(), g[tag + "/rv1"], rtol=1e-12, atol=1e-12)
        assert int(g[tag + "/nbt1"]) == 1


def test_scatter_oracle_matches_reference_module():
    g = np.load(os.path.join(GOLDEN, "pfn_small.npz"))
    y = torch.from_numpy(g["pos/y_train"]).float()
    canvas = pfn.scatter(y, torch.from_numpy(g["inds"]), 600, 600).numpy()
    want = np.zeros(tuple(g["scatter/shape"]), np.float32)
    idx = g["scatter/nz_index"]
    want[idx[:, 0], idx[:, 1], idx[:, 2], idx[:, 3]] = g["scatter/nz_value"]
    np.testing.assert_array_equal(canvas, want)
