"""oracle/aggregate.py (restated Lyft SDK transform / remove_close + dataset.py:54-88 loop): hand-checkable cases."""
import numpy as np

from oracle import aggregate as og


def test_identity_keeps_float32_values_and_drops_close_points():
    raw = np.array([[1.5, -2.25, 0.5, 7.0, 3.0],
                    [0.0005, 0.0009, 4.0, 1.0, 0.0],           # |x|, |y| < 0.001 -> removed
                    [0.0005, 0.001, 4.0, 1.0, 0.0],            # y == float32(0.001): not < radius -> kept
                    [-0.0009, 0.5, 1.0, 2.0, 0.0]], np.float32)
    pts, keeps = og.aggregate([raw], [np.eye(4)])
    assert pts.dtype == np.float64 and pts.shape == (3, 4)
    assert keeps[0].tolist() == [True, False, True, True]
    assert np.array_equal(pts, raw[[0, 2, 3], :4].astype(np.float64))


def test_transform_rounds_once_to_float32_and_concatenates_in_visit_order():
    rng = np.random.default_rng(0)
    a = rng.normal(0, 20, (50, 5)).astype(np.float32)
    b = rng.normal(0, 20, (30, 5)).astype(np.float32)
    Ma = og.pose_matrix([1.0, -2.0, 0.3], [0.9, 0.1, -0.2, 0.3])
    Mb = og.pose_matrix([4.0, 0.5, -1.0], [0.7, -0.1, 0.05, 0.7], inverse=True)
    pts, keeps = og.aggregate([a, b], [Ma, Mb])
    assert pts.shape == (80, 4) and all(k.all() for k in keeps)
    want = (Ma[:3, :3] @ a[:, :3].astype(np.float64).T + Ma[:3, 3:4]).T
    assert np.array_equal(pts[:50, :3].astype(np.float32), pts[:50, :3])            # float32-representable
    assert np.abs(pts[:50, :3] - want).max() < 4e-6 * 60                               # one float32 rounding
    assert np.array_equal(pts[:50, 3], a[:, 3].astype(np.float64))                    # intensity untouched
    want_b = (Mb[:3, :3] @ b[:, :3].astype(np.float64).T + Mb[:3, 3:4]).T
    assert np.abs(pts[50:, :3] - want_b).max() < 4e-6 * 60


def test_pose_matrix_inverse_is_the_inverse():
    M = og.pose_matrix([3.0, -1.0, 2.0], [0.3, 0.5, -0.4, 0.7])
    Mi = og.pose_matrix([3.0, -1.0, 2.0], [0.3, 0.5, -0.4, 0.7], inverse=True)
    assert np.abs(M @ Mi - np.eye(4)).max() < 1e-14
    assert abs(np.linalg.det(M[:3, :3]) - 1.0) < 1e-14
