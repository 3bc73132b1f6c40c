"""IoU oracle: closed forms, exact rational arithmetic, and the reference build."""
import numpy as np
import pytest

from oracle import native
from oracle.exact_iou import exact_iou


def rect(cx, cy, l, w, yaw=0.0):
    """CCW ring in the bottom_corners() order."""
    xs = l / 2 * np.array([1, 1, -1, -1.0])
    ys = w / 2 * np.array([-1, 1, 1, -1.0])
    c, s = np.cos(yaw), np.sin(yaw)
    return np.stack([c * xs - s * ys + cx, s * xs + c * ys + cy], 1)


def cw(ring):
    return ring[::-1].copy()


def test_identical_boxes():
    a = rect(3, 4, 5, 2, 0.3)
    assert native.iou(a, cw(a)) == pytest.approx(1.0, abs=1e-14)


def test_disjoint_and_touching():
    a = rect(0, 0, 2, 2)
    assert native.iou(a, cw(rect(5, 0, 2, 2))) == 0.0
    assert native.iou(a, cw(rect(2, 0, 2, 2))) == 0.0       # shared edge: zero area -> 0
    assert native.iou(a, cw(rect(2, 2, 2, 2))) == 0.0       # shared corner


def test_axis_aligned_overlap_closed_form():
    a = rect(0, 0, 4, 2)
    g = cw(rect(1, 0.5, 4, 2))
    inter = 3 * 1.5
    assert native.iou(a, g) == pytest.approx(inter / (8 + 8 - inter), abs=1e-15)


def test_containment():
    a = rect(0, 0, 10, 10)
    assert native.iou(a, cw(rect(1, 1, 2, 2, 0.7))) == pytest.approx(4 / 100, abs=1e-15)


def test_square_vs_45deg_copy():
    a = rect(0, 0, 2, 2)
    g = cw(rect(0, 0, 2, 2, np.pi / 4))
    inter = 8 * (np.sqrt(2) - 1)          # regular octagon of inradius 1
    assert native.iou(a, g) == pytest.approx(inter / (8 - inter), abs=1e-14)


def test_wrong_winding_is_negative():
    a = rect(0, 0, 2, 2)
    # GT passed counter-clockwise: the reference would print "IOU < 0" and exit (pillars.cpp:166-169)
    assert native.iou(a, rect(0.5, 0, 2, 2)) <= 0.0


def test_random_pairs_vs_exact_rational():
    rng = np.random.default_rng(0)
    worst = 0.0
    n_pos = 0
    for _ in range(1500):
        a = rect(rng.uniform(-3, 3), rng.uniform(-3, 3), rng.uniform(1, 12), rng.uniform(1, 6), rng.uniform(-np.pi, np.pi))
        g = cw(rect(rng.uniform(-3, 3), rng.uniform(-3, 3), rng.uniform(1, 12), rng.uniform(1, 6), rng.uniform(-np.pi, np.pi)))
        v, e = native.iou(a, g), exact_iou(a, g)
        worst = max(worst, abs(v - e))
        n_pos += e > 0
    assert n_pos > 1000
    assert worst < 1e-13


def test_make_ious_prefilter_and_reference_build(ref_module):
    rng = np.random.default_rng(1)
    A, G = 400, 7
    ac = np.stack([rng.uniform(0, 40, A), rng.uniform(0, 40, A), np.zeros(A)], 1)
    gc = np.stack([rng.uniform(0, 40, G), rng.uniform(0, 40, G), np.zeros(G)], 1)
    a_cor = np.stack([rect(ac[i, 0], ac[i, 1], 8, 4, rng.uniform(-3, 3)) for i in range(A)])
    g_cor = np.stack([cw(rect(gc[j, 0], gc[j, 1], 9, 4, rng.uniform(-3, 3))) for j in range(G)])
    ious = np.full((A, G), -1.0)
    native.make_ious(a_cor, g_cor, ac, gc, ious)
    far = (np.abs(ac[:, None, 0] - gc[None, :, 0]) > 10) | (np.abs(ac[:, None, 1] - gc[None, :, 1]) > 10)
    assert np.all(ious[far] == 0)
    assert (ious[~far] > 0).sum() > 50
    # exactly-on-the-radius pairs are NOT filtered (strict >, data/pillars.cpp:418-419)
    ac2 = np.array([[0.0, 0.0, 0.0]]); gc2 = np.array([[10.0, 0.0, 0.0]])
    o = np.zeros((1, 1))
    native.make_ious(rect(0, 0, 30, 4)[None], cw(rect(10, 0, 4, 4))[None], ac2, gc2, o)
    assert o[0, 0] > 0
    if ref_module is not None:
        r = np.full((A, G), -1.0)
        ref_module.make_ious(a_cor, g_cor, ac, gc, r)
        np.testing.assert_array_equal(ious, r)
