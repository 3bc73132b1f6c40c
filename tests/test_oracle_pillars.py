"""The C restatement (oracle/pp_oracle.c) against the reference's own data/pillars.cpp compiled
with the Boost stand-in (oracle/_ref), bit for bit, plus hand-checked branch cases."""
import numpy as np
import pytest

from oracle import native
from helpers import (GRID, cloud_boundaries, cloud_dense_cells, cloud_random, f32_exact,
                     run_create_pillars)


def test_boundaries_known_answers():
    pts = cloud_boundaries()
    t, ind = run_create_pillars(native.create_pillars, pts, 16, 4)
    # pillar 2 holds the two duplicates and (0.15,0.12,1.0) in input order, then 0.19999 (row 10)
    np.testing.assert_array_equal(ind[2], [1, 300, 299])
    np.testing.assert_array_equal(t[2, 0, :4], pts[5, :4])
    np.testing.assert_array_equal(t[2, 1, :4], pts[6, :4])
    np.testing.assert_array_equal(t[2, 2, :4], pts[7, :4])
    np.testing.assert_array_equal(t[2, 3, :4], pts[10, :4])
    # xp = canvas_x - x, yp = canvas_y - y (data/pillars.cpp:30-31)
    assert t[2, 0, 4] == 300 - pts[5, 0] and t[2, 0, 5] == 299 - pts[5, 1]
    # sequential running mean over the four points (data/pillars.cpp:324-326)
    m = pts[5, 0]
    for k, i in enumerate([6, 7, 10], start=1):
        m = m * (k / (k + 1)) + pts[i, 0] / (k + 1)
    assert t[2, 2, 6] == m - pts[7, 0]


def test_boundaries_cells():
    pts = cloud_boundaries()
    t, ind = run_create_pillars(native.create_pillars, pts, 16, 4)
    n = int(ind[:, 0].sum())
    cells = [tuple(r[1:].astype(int)) for r in ind[:n]]
    assert cells[0] == (0, 599)
    assert cells[1] == (599, 0)
    assert cells[2] == (300, 299)
    assert cells[3] == (299, 300)
    # 0.19999 -> cell 300 (same pillar as row 5), 0.2 -> cell 301; y=0.0 -> iy 300 -> cy 299
    assert (301, 299) in cells
    assert len(set(cells)) == n
    assert np.all(ind[n:] == 0)


def test_empty_and_all_out_of_range():
    t, ind = run_create_pillars(native.create_pillars, np.zeros((0, 4)), 8, 4)
    assert not t.any() and not ind.any()
    pts = f32_exact([[100, 0, 0, 1], [0, -100, 0, 1], [0, 0, 11, 1]])
    t, ind = run_create_pillars(native.create_pillars, pts, 8, 4)
    assert not t.any() and not ind.any()


def test_first_n_cap_and_mean_over_all_points():
    pts = cloud_dense_cells(3, n=3000, ncells=5)
    N = 50
    t, ind = run_create_pillars(native.create_pillars, pts, 16, N)
    assert int(ind[:, 0].sum()) == 5
    cx = np.floor((pts[:, 0] + 60) / .2)
    cy = 599 - np.floor((pts[:, 1] + 60) / .2)
    for p in range(5):
        sel = np.where((cx == ind[p, 1]) & (cy == ind[p, 2]))[0]
        assert len(sel) > N
        np.testing.assert_array_equal(t[p, :, :4], pts[sel[:N], :4])    # first N in input order
        m = pts[sel[0], 0]
        for k, i in enumerate(sel[1:], start=1):                        # mean over ALL points
            m = m * (k / (k + 1)) + pts[i, 0] / (k + 1)
        np.testing.assert_array_equal(t[p, :, 6], m - pts[sel[:N], 0])


def test_pillar_cap_keeps_first_touched():
    pts = cloud_random(5, n=4000, spread=59.0)
    t_all, ind_all = run_create_pillars(native.create_pillars, pts, 4000, 8)
    n_all = int(ind_all[:, 0].sum())
    P = n_all // 3
    t, ind = run_create_pillars(native.create_pillars, pts, P, 8)
    assert int(ind[:, 0].sum()) == P
    np.testing.assert_array_equal(ind, ind_all[:P])
    np.testing.assert_array_equal(t, t_all[:P])


@pytest.mark.parametrize("case", ["boundaries", "random", "dense", "strided", "capP"])
def test_matches_reference_build(ref_module, case):
    if ref_module is None:
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    P, N = 3000, 20
    if case == "boundaries":
        pts = cloud_boundaries()
    elif case == "random":
        pts = cloud_random(11, n=6000)
    elif case == "dense":
        pts = cloud_dense_cells(12)
    elif case == "strided":
        pts = np.ascontiguousarray(cloud_random(13, n=3000, cols=4).T).T   # the reference's [4,N].T view
        assert not pts.flags.c_contiguous
    else:
        pts = cloud_random(14, n=6000, spread=59.0)
        P = 500
    t0, i0 = run_create_pillars(ref_module.create_pillars, pts, P, N)
    t1, i1 = run_create_pillars(native.create_pillars, pts, P, N)
    np.testing.assert_array_equal(i1, i0)
    np.testing.assert_array_equal(t1, t0)     # bit for bit, including the running means
