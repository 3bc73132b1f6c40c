"""Shared test inputs: hand-built clouds hitting each branch of create_pillars and seeded sweeps."""
import numpy as np

GRID = (.2, .2, -60, -60, -10, 60, 60, 10, 600)   # x_step,y_step,x_min,y_min,z_min,x_max,y_max,z_max,H


def f32_exact(a):
    """float64 array whose values are exactly float32-representable."""
    return np.asarray(a, dtype=np.float32).astype(np.float64)


def cloud_boundaries():
    # x == x_max excluded, x == x_min included, same for y/z; a far-outside point; duplicates
    pts = [
        [-60.0, -60.0, -10.0, 1.0],    # all mins: included, cell (0, 599)
        [60.0, 0.0, 0.0, 2.0],         # x == x_max: excluded
        [0.0, 60.0, 0.0, 3.0],         # y == y_max: excluded
        [0.0, 0.0, 10.0, 4.0],         # z == z_max: excluded
        [59.99, 59.99, 9.99, 5.0],     # just inside the max corner
        [0.1, 0.1, 0.0, 6.0], [0.1, 0.1, 0.0, 6.0], [0.15, 0.12, 1.0, 7.0],   # duplicates + same cell
        [-0.1, -0.1, 0.0, 8.0],
        [1000.0, 0.0, 0.0, 9.0],
        [0.19999, 0.0, 0.0, 10.0], [0.2, 0.0, 0.0, 11.0],   # either side of a cell edge
    ]
    return f32_exact(pts)


def cloud_random(seed, n=5000, spread=70.0, cols=5):
    rng = np.random.default_rng(seed)
    pts = rng.uniform(-spread, spread, (n, cols))
    pts[:, 2] = rng.uniform(-12, 12, n)
    return f32_exact(pts)


def cloud_dense_cells(seed, n=6000, ncells=12):
    """Many points in few cells: exercises the first-N cap and long running means."""
    rng = np.random.default_rng(seed)
    cx = rng.integers(0, 600, ncells)
    cy = rng.integers(0, 600, ncells)
    which = rng.integers(0, ncells, n)
    x = -60 + (cx[which] + rng.uniform(0.05, 0.95, n)) * 0.2
    y = -60 + (cy[which] + rng.uniform(0.05, 0.95, n)) * 0.2
    z = rng.uniform(-3, 3, n)
    r = rng.uniform(0, 255, n)
    return f32_exact(np.stack([x, y, z, r], 1))


def run_create_pillars(fn, pts, P, N, grid=GRID):
    t = np.zeros((P, N, 9))
    ind = np.zeros((P, 3))
    fn(pts, t, ind, N, P, *grid)
    return t, ind


def boxes_from_gt(gt, Box, names):
    return [Box(gt["centers"][i], gt["wlh"][i], gt["yaw"][i], names[int(gt["cls"][i])])
            for i in range(len(gt["yaw"]))]


def targets_fixture():
    """tests/golden/targets_small.npz (made by the reference's own utils/box_utils.py, see
    tests/golden/make_golden_targets.py) as a dict, plus dense (cls, reg, ious) builders."""
    import os
    here = os.path.dirname(os.path.abspath(__file__))
    z = np.load(os.path.join(here, "golden", "targets_small.npz"))
    return {k: z[k] for k in z.files}


def fixture_case(fx, tag):
    """Dense reference outputs of one fixture case: dict with the GT arrays and cls [A,9], reg [A,9], ious [A,G]."""
    A = fx["a_centers"].shape[0]
    G = fx[tag + "/g_cls"].shape[0]
    cls = np.zeros((A, fx[tag + "/cls_rows"].shape[1])); reg = np.zeros((A, fx[tag + "/reg_rows"].shape[1]))
    cls[fx[tag + "/rows"]] = fx[tag + "/cls_rows"]
    reg[fx[tag + "/rows"]] = fx[tag + "/reg_rows"]
    ious = np.zeros((A, G))
    ious[fx[tag + "/iou_a"], fx[tag + "/iou_g"]] = fx[tag + "/iou_v"]
    out = {k.split("/", 1)[1]: v for k, v in fx.items() if k.startswith(tag + "/")}
    out.update(cls=cls, reg=reg, ious=ious)
    return out
