"""Stand-in for ``lyft_dataset_sdk.utils.geometry_utils`` (absent from this image): the two functions the
reference imports (utils/box_utils.py:15,17; data/dataset.py).  Published behaviour restated from memory of the
SDK source (a fork of nuscenes-devkit) -- UNPINNED.  TEST INFRASTRUCTURE ONLY."""
import numpy as np
from pyquaternion import Quaternion


def transform_matrix(translation=np.array([0, 0, 0]), rotation=Quaternion([1, 0, 0, 0]), inverse=False):
    """4x4 homogeneous transform [R | t] (or its inverse [R^T | -R^T t])."""
    tm = np.eye(4)
    if inverse:
        rot_inv = rotation.rotation_matrix.T
        trans = np.transpose(-np.array(translation))
        tm[:3, :3] = rot_inv
        tm[:3, 3] = rot_inv.dot(trans)
    else:
        tm[:3, :3] = rotation.rotation_matrix
        tm[:3, 3] = np.transpose(np.array(translation))
    return tm


def points_in_box(box, points, wlh_factor=1.0):
    """Mask [n] of the columns of ``points`` [3,n] inside ``box`` (projection on the three box edges
    that meet at corner 0)."""
    corners = box.corners(wlh_factor=wlh_factor)
    p1 = corners[:, 0]
    p_x = corners[:, 4]
    p_y = corners[:, 1]
    p_z = corners[:, 3]
    i = p_x - p1
    j = p_y - p1
    k = p_z - p1
    v = points - p1.reshape((-1, 1))
    iv = np.dot(i, v)
    jv = np.dot(j, v)
    kv = np.dot(k, v)
    mask_x = np.logical_and(0 <= iv, iv <= np.dot(i, i))
    mask_y = np.logical_and(0 <= jv, jv <= np.dot(j, j))
    mask_z = np.logical_and(0 <= kv, kv <= np.dot(k, k))
    return np.logical_and(np.logical_and(mask_x, mask_y), mask_z)
