"""Stand-in for the ``pyquaternion`` package (absent from this image; a transitive, unpinned dependency of
the reference, call sites utils/box_utils.py:80-81,147,246,288).  TEST INFRASTRUCTURE ONLY.

Restates the published behaviour of ``pyquaternion.Quaternion`` for the members the reference reaches:
construction from ``axis=``/``degrees=``/``radians=``, from a 4-sequence or four scalars, ``rotation_matrix``,
``yaw_pitch_roll``, ``inverse``, ``rotate``, multiplication and iteration.  Formulas quoted from memory of the
package source (UNPINNED, see README.md): q = (w, x, y, z);

    from axis/angle   q = (cos(a/2), axis_unit * sin(a/2)),  degrees -> float(d) / 180.0 * pi
    rotation_matrix   normalise, then (Q . conj(Qbar)^T)[1:, 1:] with the 4x4 left / right product matrices
    yaw_pitch_roll    normalise, then yaw = atan2(2(w z - x y), 1 - 2(y^2 + z^2)),
                      pitch = asin(2(w y + z x)), roll = atan2(2(w x - y z), 1 - 2(x^2 + y^2))
"""
from math import cos, pi, sin, sqrt

import numpy as np


class Quaternion:
    def __init__(self, *args, **kwargs):
        if len(args) == 0:
            if "axis" in kwargs or "radians" in kwargs or "degrees" in kwargs or "angle" in kwargs:
                axis = self._seq(kwargs["axis"], 3)
                angle = kwargs.get("radians") or self.to_radians(kwargs.get("degrees")) or kwargs.get("angle") or 0.0
                self.q = Quaternion._from_axis_angle(axis, angle).q
            elif "array" in kwargs:
                self.q = self._seq(kwargs["array"], 4)
            elif "scalar" in kwargs or "vector" in kwargs:
                v = kwargs.get("vector")
                self.q = np.hstack(([float(kwargs.get("scalar") or 0.0)], self._seq(v, 3) if v is not None else np.zeros(3)))
            elif kwargs:
                raise ValueError("sdk_shim.Quaternion: unsupported keyword form %r" % sorted(kwargs))
            else:
                self.q = np.array([1.0, 0.0, 0.0, 0.0])
        elif len(args) == 1:
            if isinstance(args[0], Quaternion):
                self.q = args[0].q
                return
            if args[0] is None:
                raise TypeError("Object cannot be initialised from None")
            try:
                r = float(args[0])
                self.q = np.zeros(4)
                self.q[0] = r
                return
            except TypeError:
                pass
            self.q = self._seq(args[0], 4)
        else:
            self.q = self._seq(args, 4)

    @staticmethod
    def _seq(seq, n):
        if len(seq) != n:
            raise ValueError("Unexpected number of elements in sequence. Got: %d, Expected: %d." % (len(seq), n))
        return np.array([float(e) for e in seq])

    @classmethod
    def to_radians(cls, angle_deg):
        if angle_deg is not None:
            return float(angle_deg) / 180.0 * pi

    @classmethod
    def _from_axis_angle(cls, axis, angle):
        mag_sq = np.dot(axis, axis)
        if mag_sq == 0.0:
            raise ZeroDivisionError("Provided rotation axis has no length")
        if abs(1.0 - mag_sq) > 1e-12:
            axis = axis / sqrt(mag_sq)
        theta = angle / 2.0
        r = cos(theta)
        i = axis * sin(theta)
        return cls(r, i[0], i[1], i[2])

    # --- norms -------------------------------------------------------------------------------
    def _sum_of_squares(self):
        return np.dot(self.q, self.q)

    @property
    def norm(self):
        return sqrt(self._sum_of_squares())

    def is_unit(self, tolerance=1e-14):
        return abs(1.0 - self._sum_of_squares()) < tolerance

    def _normalise(self):
        if not self.is_unit():
            n = self.norm
            if n > 0:
                self.q = self.q / n

    # --- algebra -----------------------------------------------------------------------------
    def _q_matrix(self):
        q = self.q
        return np.array([[q[0], -q[1], -q[2], -q[3]],
                         [q[1], q[0], -q[3], q[2]],
                         [q[2], q[3], q[0], -q[1]],
                         [q[3], -q[2], q[1], q[0]]])

    def _q_bar_matrix(self):
        q = self.q
        return np.array([[q[0], -q[1], -q[2], -q[3]],
                         [q[1], q[0], q[3], -q[2]],
                         [q[2], -q[3], q[0], q[1]],
                         [q[3], q[2], -q[1], q[0]]])

    def __mul__(self, other):
        if isinstance(other, Quaternion):
            return Quaternion(array=np.dot(self._q_matrix(), other.q))
        return self * Quaternion(other)

    @property
    def conjugate(self):
        return Quaternion(scalar=self.q[0], vector=-self.q[1:4])

    @property
    def inverse(self):
        ss = self._sum_of_squares()
        if ss > 0:
            return Quaternion(array=(self._vector_conjugate() / ss))
        raise ZeroDivisionError("a zero quaternion (0 + 0i + 0j + 0k) cannot be inverted")

    def _vector_conjugate(self):
        return np.hstack((self.q[0], -self.q[1:4]))

    @property
    def rotation_matrix(self):
        self._normalise()
        product_matrix = np.dot(self._q_matrix(), self._q_bar_matrix().conj().transpose())
        return product_matrix[1:][:, 1:]

    def rotate(self, vector):
        if isinstance(vector, Quaternion):
            return self._rotate_quaternion(vector)
        q = Quaternion(vector=vector)
        a = self._rotate_quaternion(q).q[1:4]
        if isinstance(vector, list):
            return [float(e) for e in a]
        if isinstance(vector, tuple):
            return tuple(float(e) for e in a)
        return a

    def _rotate_quaternion(self, q):
        self._normalise()
        return self * q * self.conjugate

    @property
    def yaw_pitch_roll(self):
        self._normalise()
        q = self.q
        yaw = np.arctan2(2 * (q[0] * q[3] - q[1] * q[2]), 1 - 2 * (q[2] ** 2 + q[3] ** 2))
        pitch = np.arcsin(2 * (q[0] * q[2] + q[3] * q[1]))
        roll = np.arctan2(2 * (q[0] * q[1] - q[2] * q[3]), 1 - 2 * (q[1] ** 2 + q[2] ** 2))
        return yaw, pitch, roll

    # --- container protocol (``Quaternion(list(box.orientation))``, utils/box_utils.py:288) ----
    def __iter__(self):
        return iter(self.q)

    def __len__(self):
        return 4

    def __getitem__(self, i):
        return self.q[int(i)]

    @property
    def elements(self):
        return self.q

    def __repr__(self):
        return "Quaternion({!r}, {!r}, {!r}, {!r})".format(*[float(e) for e in self.q])
