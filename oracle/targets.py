"""numpy restatement of /root/reference/utils/box_utils.py target assignment.
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

PINNED: ``create_target`` / ``make_target`` / ``boxes_to_image_space`` / the anchor lattice reproduce, bit for
bit, the outputs of the reference's own utils/box_utils.py run unmodified in the build container
(oracle/refpy.py + oracle/sdk_shim; fixture tests/golden/targets_small.npz made by
tests/golden/make_golden_targets.py; tests/test_oracle_targets_pinned.py, which also re-runs the reference live
where /root/reference exists).  Still from memory of the absent packages (see oracle/sdk_shim/README.md): the
arithmetic inside lyft_dataset_sdk ``Box.corners()`` and pyquaternion.

``Box`` below is a duck-typed stand-in exposing exactly the attributes the reference touches: ``center``,
``wlh``, ``name``, ``orientation.yaw_pitch_roll[0]`` and ``bottom_corners()`` (corner order [2,3,7,6] of the
nuScenes-style corner table).  It rotates with cos/sin of the yaw; the SDK goes through a quaternion rotation
matrix and ``np.dot`` (BLAS, FMA-dependent last bits), so corners agree to 1 ulp, not bit for bit.
"""
import numpy as np

from . import config as cfg
from . import native


class _Orientation:
    def __init__(self, yaw):
        self.yaw_pitch_roll = (float(yaw), 0.0, 0.0)


class Box:
    """Stand-in for lyft_dataset_sdk.utils.data_classes.Box restricted to yaw-only rotation."""

    def __init__(self, center, size, yaw, name=None):
        self.center = np.array(center, dtype=np.float64)
        self.wlh = np.array(size, dtype=np.float64)
        self.orientation = _Orientation(yaw)
        self.name = name

    @property
    def yaw(self):
        return self.orientation.yaw_pitch_roll[0]

    def bottom_corners(self):
        """[3,4]: (+l/2,-w/2), (+l/2,+w/2), (-l/2,+w/2), (-l/2,-w/2) rotated by yaw + center."""
        w, l, h = self.wlh
        xs = l / 2 * np.array([1, 1, -1, -1.0])
        ys = w / 2 * np.array([-1, 1, 1, -1.0])
        zs = h / 2 * np.array([-1, -1, -1, -1.0])
        c, s = np.cos(self.yaw), np.sin(self.yaw)
        out = np.vstack((c * xs - s * ys, s * xs + c * ys, zs))
        return out + self.center.reshape(3, 1)


def _sdk_yaw(deg):
    """yaw_pitch_roll[0] of pyquaternion's Quaternion(axis=[0,0,1], degrees=deg) (utils/box_utils.py:80,147):
    atan2(2 w z, 1 - 2 z^2) with (w, z) = (cos(a/2), sin(a/2)), a = deg/180*pi.  90 degrees -> pi/2 - 2.2e-16.
    Pinned against the reference run through oracle/sdk_shim (tests/test_oracle_targets_pinned.py)."""
    a = float(deg) / 180.0 * np.pi
    w, z = np.cos(a / 2.0), np.sin(a / 2.0)
    return float(np.arctan2(2 * (w * z), 1 - 2 * (z ** 2)))


def boxes_to_image_space(boxes):
    """utils/box_utils.py:19-32 -- centers [G,3] and bottom corners [G,4,2], y flipped."""
    centers = np.stack([box.center.copy() for box in boxes])
    corners = np.stack([box.bottom_corners().transpose([1, 0])[:, :2] for box in boxes])
    centers[..., 1] = (cfg.CANVAS_HEIGHT - 1) - centers[..., 1]
    corners[..., 1] = (cfg.CANVAS_HEIGHT - 1) - corners[..., 1]
    return centers, corners


def make_target(anchor_box, gt_box, anch=None):
    """utils/box_utils.py:70-109."""
    ax, ay, az = anchor_box.center
    gx, gy, gz = gt_box.center
    aw, al, ah = anchor_box.wlh
    gw, gl, gh = gt_box.wlh
    ad = np.sqrt(aw ** 2 + al ** 2)
    at = anchor_box.orientation.yaw_pitch_roll[0]
    gt = gt_box.orientation.yaw_pitch_roll[0]

    gy = (cfg.CANVAS_HEIGHT - 1) - gy
    dx = (gx - ax) / ad
    dy = (gy - ay) / ad
    dz = (gz - az) / ah

    dw = np.log(gw / aw)
    dl = np.log(gl / al)
    dh = np.log(gh / ah)

    if (gt <= np.pi and gt >= np.pi / 2):
        gt -= np.pi
    elif (gt >= -np.pi and gt <= -np.pi / 2):
        gt += np.pi

    dt = np.sin(gt - at)

    if ((gt - at) <= np.pi and (gt - at) >= np.pi / 2) or ((gt - at) >= -np.pi and (gt - at) <= -np.pi / 2):
        ort = 1
    else:
        ort = 0
    return [1, dx, dy, dz, dw, dl, dh, dt, ort]


def make_anchor_boxes(fm_height=None, fm_width=None):
    """utils/box_utils.py:111-159 -- returns (boxes, corners [A,4,2], centers [A,3], xy [A,4]).
    Anchor index a = (y*fm_width + x)*6 + d."""
    fm_height = int(cfg.FM_HEIGHT if fm_height is None else fm_height)
    fm_width = int(cfg.FM_WIDTH if fm_width is None else fm_width)
    corners_list, boxes_list, centers_list, xy_list = [], [], [], []
    for y in range(fm_height):
        for x in range(fm_width):
            for d in range(len(cfg.ANCHOR_DIMS)):
                x_center = (x + 0.5) / cfg.FM_SCALE
                y_center = (y + 0.5) / cfg.FM_SCALE
                z_center = cfg.ANCHOR_ZS[d]
                width, length, height = cfg.ANCHOR_DIMS[d]
                yaw = cfg.ANCHOR_YAWS[d]
                box = Box([x_center, y_center, z_center], [width, length, height], _sdk_yaw(yaw))
                boxes_list.append(box)
                bc = box.bottom_corners().transpose([1, 0])
                corners_list.append(bc[:, :2])
                centers_list.append([x_center, y_center, z_center])
                if yaw > 0:
                    xy_list.append(np.concatenate((bc[1, :2], bc[3, :2])))
                else:
                    xy_list.append(np.concatenate((bc[2, :2], bc[0, :2])))
    return boxes_list, np.array(corners_list), np.array(centers_list), np.array(xy_list)


def create_target(anchor_corners, gt_corners, anchor_centers, gt_centers, anchor_box_list,
                  gt_box_list, make_ious=None, return_ious=False):
    """utils/box_utils.py:162-232, statement by statement.  ``make_ious`` defaults to the C
    oracle (oracle/pp_oracle.c); pass ``oracle.ref.load().make_ious`` to run the reference's."""
    pos_thresh = cfg.IOU_POS_THRESH
    make_ious = native.make_ious if make_ious is None else make_ious

    ious = np.zeros((len(anchor_box_list), len(gt_box_list)))
    make_ious(anchor_corners, gt_corners, anchor_centers, gt_centers, ious)
    ious_dense = ious

    cls_targets = np.zeros((len(anchor_box_list), cfg.NUM_CLASSES))
    reg_targets = np.zeros((len(anchor_box_list), cfg.REG_DIMS + 1))

    gt_box_classes = np.array([cfg.NAME_TO_IND[box.name] for box in gt_box_list], dtype=np.int32)

    max_ious = np.max(ious, axis=1)
    arg_max_ious = np.argmax(ious, axis=1)
    pos_anchors = np.where(max_ious > pos_thresh)[0]
    pos_boxes = arg_max_ious[pos_anchors]

    ious = ious.transpose([1, 0])
    top_anchor_for_box = np.argmax(ious, axis=1)

    filter_inds = np.nonzero(top_anchor_for_box)
    top_anchor_for_box = top_anchor_for_box[filter_inds]

    cls_targets[pos_anchors, gt_box_classes[pos_boxes]] = 1
    cls_targets[top_anchor_for_box, :] = 0
    cls_targets[top_anchor_for_box, gt_box_classes[filter_inds]] = 1

    for i, anch in enumerate(pos_anchors):
        reg_targets[anch, :] = make_target(anchor_box_list[anch], gt_box_list[pos_boxes[i]], anch)

    matched_boxes = [gt_box_list[i] for i in filter_inds[0]]
    for i, anch in enumerate(top_anchor_for_box):
        reg_targets[anch, :] = make_target(anchor_box_list[anch], matched_boxes[i], anch)

    if return_ious:
        return cls_targets, reg_targets, ious_dense
    return cls_targets, reg_targets


class LazyAnchorBoxes:
    """Indexable stand-in for the reference's 540000-element ``anchor_box_list``: create_target
    only indexes it for positive / forced anchors (utils/box_utils.py:219-228)."""

    def __init__(self, centers, wlh, yaw):
        self.centers, self.wlh, self.yaw = centers, wlh, yaw

    def __len__(self):
        return self.centers.shape[0]

    def __getitem__(self, a):
        return Box(self.centers[a], self.wlh[a], self.yaw[a])


def anchor_arrays(fm_height=None, fm_width=None):
    """Vectorised equivalent of make_anchor_boxes() (checked against the loop in
    tests/test_oracle_targets.py): corners [A,4,2], centers [A,3], wlh [A,3], yaw [A]."""
    fm_height = int(cfg.FM_HEIGHT if fm_height is None else fm_height)
    fm_width = int(cfg.FM_WIDTH if fm_width is None else fm_width)
    nd = len(cfg.ANCHOR_DIMS)
    ys, xs, ds = np.meshgrid(np.arange(fm_height), np.arange(fm_width), np.arange(nd), indexing="ij")
    ys, xs, ds = ys.reshape(-1), xs.reshape(-1), ds.reshape(-1)
    wlh = np.stack(cfg.ANCHOR_DIMS)[ds]
    yaw = np.array([_sdk_yaw(d) for d in cfg.ANCHOR_YAWS], dtype=np.float64)[ds]
    centers = np.stack([(xs + 0.5) / cfg.FM_SCALE, (ys + 0.5) / cfg.FM_SCALE,
                        np.asarray(cfg.ANCHOR_ZS, dtype=np.float64)[ds]], axis=1)
    w, l = wlh[:, 0:1], wlh[:, 1:2]
    bx = l / 2 * np.array([[1, 1, -1, -1.0]])
    by = w / 2 * np.array([[-1, 1, 1, -1.0]])
    c, s = np.cos(yaw)[:, None], np.sin(yaw)[:, None]
    corners = np.stack([c * bx - s * by + centers[:, 0:1], s * bx + c * by + centers[:, 1:2]], axis=2)
    return corners, centers, wlh, yaw
