"""Stand-in for ``lyft_dataset_sdk.utils.data_classes`` (absent from this image): ``Box`` and
``LidarPointCloud`` restricted to the members the reference's hot path reaches (utils/box_utils.py:26-27,
74-81,148-151,250-291; data/dataset.py:54-88).  Published behaviour restated from memory of the SDK source (a
fork of nuscenes-devkit's data_classes.py) -- UNPINNED, see ../../README.md.  TEST INFRASTRUCTURE ONLY.

Box.corners(): x forward (length), y left (width), z up; corner table
    x = l/2 * [ 1,  1,  1,  1, -1, -1, -1, -1]
    y = w/2 * [ 1, -1, -1,  1,  1, -1, -1,  1]
    z = h/2 * [ 1,  1, -1, -1,  1,  1, -1, -1]
rotated by ``orientation.rotation_matrix`` (np.dot), then translated by the centre;
bottom_corners() = corners()[:, [2, 3, 7, 6]] = (+l/2,-w/2), (+l/2,+w/2), (-l/2,+w/2), (-l/2,-w/2).
"""
import numpy as np
from pyquaternion import Quaternion


class Box:
    def __init__(self, center, size, orientation, label=np.nan, score=np.nan, velocity=(np.nan, np.nan, np.nan),
                 name=None, token=None):
        assert not np.any(np.isnan(center))
        assert not np.any(np.isnan(size))
        assert len(center) == 3
        assert len(size) == 3
        assert type(orientation) == Quaternion
        self.center = np.array(center)
        self.wlh = np.array(size)
        self.orientation = orientation
        self.label = int(label) if not np.isnan(label) else label
        self.score = float(score) if not np.isnan(score) else score
        self.velocity = np.array(velocity)
        self.name = name
        self.token = token

    @property
    def rotation_matrix(self):
        return self.orientation.rotation_matrix

    def translate(self, x):
        self.center += x

    def rotate(self, quaternion):
        self.center = np.dot(quaternion.rotation_matrix, self.center)
        self.orientation = quaternion * self.orientation
        self.velocity = np.dot(quaternion.rotation_matrix, self.velocity)

    def corners(self, wlh_factor=1.0):
        width, length, height = self.wlh * wlh_factor
        x_corners = length / 2 * np.array([1, 1, 1, 1, -1, -1, -1, -1])
        y_corners = width / 2 * np.array([1, -1, -1, 1, 1, -1, -1, 1])
        z_corners = height / 2 * np.array([1, 1, -1, -1, 1, 1, -1, -1])
        corners = np.vstack((x_corners, y_corners, z_corners))
        corners = np.dot(self.orientation.rotation_matrix, corners)
        x, y, z = self.center
        corners[0, :] = corners[0, :] + x
        corners[1, :] = corners[1, :] + y
        corners[2, :] = corners[2, :] + z
        return corners

    def bottom_corners(self):
        return self.corners()[:, [2, 3, 7, 6]]


class LidarPointCloud:
    """points: float32 [4, n] (x, y, z, intensity)."""

    def __init__(self, points):
        assert points.shape[0] == 4
        self.points = points

    @classmethod
    def from_array(cls, raw_f32_rows):
        """from_file() minus the file: rows [n, 5] float32 -> points [4, n] float32."""
        scan = np.asarray(raw_f32_rows, dtype=np.float32)
        return cls(scan.reshape((-1, 5))[:, :4].T.copy())

    @classmethod
    def from_file(cls, file_name):
        scan = np.fromfile(str(file_name), dtype=np.float32)
        return cls(scan.reshape((-1, 5))[:, :4].T)

    def nbr_points(self):
        return self.points.shape[1]

    def transform(self, transf_matrix):
        self.points[:3, :] = transf_matrix.dot(np.vstack((self.points[:3, :], np.ones(self.nbr_points()))))[:3, :]

    def remove_close(self, radius):
        x_filt = np.abs(self.points[0, :]) < radius
        y_filt = np.abs(self.points[1, :]) < radius
        not_close = np.logical_not(np.logical_and(x_filt, y_filt))
        self.points = self.points[:, not_close]
