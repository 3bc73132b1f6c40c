"""Restatement of the dataset glue around create_pillars, /root/reference/data/dataset.py:88-106.
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py)."""
import numpy as np
import torch

from . import config as cfg
from . import native


def pillarize(lidar_points, data_mean=None, max_pillars=None, max_points=None, create_pillars=None,
              grid=None):
    """lidar_points: float64 [Npts, >=4] (any strides).  Returns (pillar f32 [9,P,N] torch,
    indices int64 [P,3] torch) exactly as PPDataset.__getitem__ builds them.

    data/dataset.py:89-90  np.zeros tensor / indices (float64)
    data/dataset.py:92-97  create_pillars(...)
    data/dataset.py:99-105 transpose -> float32 -> flat subtract of the per-slot data_mean
    data/dataset.py:106    indices -> int64
    """
    P = int(cfg.MAX_PILLARS if max_pillars is None else max_pillars)
    N = int(cfg.MAX_POINTS_PER_PILLAR if max_points is None else max_points)
    create_pillars = native.create_pillars if create_pillars is None else create_pillars
    g = grid or (cfg.X_STEP, cfg.Y_STEP, cfg.X_MIN, cfg.Y_MIN, cfg.Z_MIN, cfg.X_MAX, cfg.Y_MAX,
                 cfg.Z_MAX, cfg.CANVAS_HEIGHT)
    pillar = np.zeros((P, N, 9))
    indices = np.zeros((P, 3))
    create_pillars(lidar_points, pillar, indices, N, P, *g)
    pillar = pillar.transpose([2, 0, 1])
    pillar_size = pillar.shape
    pillar = torch.from_numpy(pillar).float()
    if data_mean is not None:
        pillar = pillar.reshape(-1) - data_mean
        pillar = pillar.reshape(pillar_size)
    indices = torch.from_numpy(indices).long()
    return pillar, indices
