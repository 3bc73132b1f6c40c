"""CPU restatement of the multi-sweep aggregation in front of create_pillars
(/root/reference/data/dataset.py:54-88).  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The arithmetic lives in a third-party dependency that is NOT vendored under /root/reference:
``lyft_dataset_sdk`` (pip, unpinned -- install_mods.sh:4), ``lyft_dataset_sdk.utils.data_classes.
LidarPointCloud`` (a fork of nuscenes-devkit's class).  Its published algorithm, restated here:

  from_file      scan = np.fromfile(path, float32); points = scan.reshape(-1, 5)[:, :4].T      -> float32 [4, n]
  transform(M)   points[:3, :] = M.dot(np.vstack((points[:3, :], np.ones(n))))[:3, :]          (float64 product,
                 stored back into the float32 array: ONE rounding to float32 per coordinate)
  remove_close   keep = ~((|points[0]| < radius) & (|points[1]| < radius))                      (float32 compare)

and dataset.py: ``agg_pc = np.hstack((agg_pc, curr_pc.points))`` starting from ``np.zeros((4, 0))`` -> float64
[4, sum n], then ``agg_pc.transpose([1, 0])`` is what create_pillars receives.  ``transmat`` is the float64
product ref_car_from_global . global_from_curr_car . curr_car_from_curr_sensor (dataset.py:78), composed on the
host in both implementations.

PARITY UNPINNED for the last bit of the float64 4-term dot product: numpy hands it to the BLAS it was built
with (summation order / FMA use unknown).  This restatement uses the sequential order k = 0..3 with separately
rounded IEEE operations; the following rounding to float32 hides the difference except when the float64 value
lies within ~1e-16 relative of a float32 rounding boundary (probability ~2^-29 per coordinate).  No golden
vectors exist in the reference for this path and the SDK is absent from this image."""
import numpy as np


def transform_points(points_f32, transmat):
    """LidarPointCloud.transform on a float32 [4, n] array; returns the float32 [3, n] coordinates."""
    M = np.asarray(transmat, np.float64)
    p = np.asarray(points_f32, np.float32)
    x, y, z = (p[k].astype(np.float64) for k in range(3))
    out = np.empty((3, p.shape[1]), np.float32)
    for r in range(3):
        out[r] = (((M[r, 0] * x + M[r, 1] * y) + M[r, 2] * z) + M[r, 3] * 1.0).astype(np.float32)
    return out


def aggregate(raw_files, transmats, min_dist=0.001):
    """dataset.py:54-88 for one sample: raw_files = list of float32 [n_i, 5] arrays in the order the loop visits
    them (current sweep first, then prev_token ...), transmats = the matching 4x4 float64 matrices.
    Returns (lidar_points float64 [T', 4] -- the create_pillars argument --, keep masks per file)."""
    agg = np.zeros((4, 0))
    keeps = []
    for raw, M in zip(raw_files, transmats):
        pts = np.ascontiguousarray(np.asarray(raw, np.float32).reshape(-1, 5)[:, :4].T)      # from_file
        pts[:3, :] = transform_points(pts, M)                                                   # transform
        radius = np.float32(min_dist)                                                           # remove_close
        keep = ~((np.abs(pts[0]) < radius) & (np.abs(pts[1]) < radius))
        keeps.append(keep)
        agg = np.hstack((agg, pts[:, keep]))
    return agg.transpose([1, 0]), keeps


def pose_matrix(translation, rotation_wxyz, inverse=False):
    """lyft_dataset_sdk.utils.geometry_utils.transform_matrix with a unit quaternion (w, x, y, z) (pyquaternion's
    rotation_matrix), float64: [[R, t], [0, 1]] or its inverse [[R^T, -R^T t], [0, 1]]."""
    w, x, y, z = (float(v) for v in rotation_wxyz)
    n = np.sqrt(w * w + x * x + y * y + z * z)
    w, x, y, z = w / n, x / n, y / n, z / n
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                  [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                  [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
    t = np.asarray(translation, np.float64)
    M = np.eye(4)
    if inverse:
        M[:3, :3] = R.T
        M[:3, 3] = R.T.dot(-t)
    else:
        M[:3, :3] = R
        M[:3, 3] = t
    return M
