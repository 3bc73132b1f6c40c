"""Grid / tensor-shape parameters of the hot path.

Defaults are the values the reference's config.py evaluates to (config.py:46-61,97,119-122,
133-134); every one is a runtime argument of the kernels.
"""
from dataclasses import dataclass, field

import numpy as np

from ._lib import PPGrid


@dataclass
class PPConfig:
    x_min: float = -60.0          # config.py:46
    y_min: float = -60.0          # config.py:47
    z_min: float = -10.0          # config.py:48
    x_max: float = 60.0           # config.py:49
    y_max: float = 60.0           # config.py:50
    z_max: float = 10.0           # config.py:51
    x_step: float = 0.2           # config.py:52
    y_step: float = 0.2           # config.py:53
    fm_scale: float = 0.5         # config.py:55
    canvas_height: int = 600      # config.py:60
    canvas_width: int = 600       # config.py:61
    fm_height: int = 300          # config.py:58
    fm_width: int = 300           # config.py:59
    max_points_per_pillar: int = 200   # config.py:119
    max_pillars: int = 24000           # config.py:120
    reg_dims: int = 8                  # config.py:121
    iou_pos_thresh: float = 0.6        # config.py:122
    num_classes: int = 9               # config.py:97
    feature_net_in: int = 9            # config.py:133
    feature_net_out: int = 64          # config.py:134
    class_names: tuple = ('animal', 'bicycle', 'bus', 'car', 'emergency_vehicle', 'motorcycle',
                          'other_vehicle', 'pedestrian', 'truck')  # config.py:98-99
    anchor_yaws_deg: tuple = (0, 90, 0, 90, 0, 90)                 # config.py:113
    anchor_zs: tuple = (.5, .5, .75, .75, 1.0, 1.0)                # config.py:115
    anchor_dims: tuple = field(default_factory=lambda: _default_anchor_dims())  # config.py:64-89,109

    def grid(self):
        return PPGrid(self.x_step, self.y_step, self.x_min, self.y_min, self.z_min, self.x_max,
                      self.y_max, self.z_max, float(self.canvas_height))

    @property
    def name_to_ind(self):
        return {n: i for i, n in enumerate(self.class_names)}   # config.py:128


def class_dims(step=0.2):
    """Approximate per-class (w, l, h) with w,l in canvas units: config.py:64-81."""
    raw = {'animal': (.5, 1, .5), 'bicycle': (.75, 2, 1.5), 'bus': (3, 12.5, 3.5),
           'car': (2, 5, 1.75), 'emergency_vehicle': (2.5, 6.5, 2.5), 'motorcycle': (1, 2.5, 1.5),
           'other_vehicle': (2.75, 8.5, 3.5), 'pedestrian': (.75, .75, 1.75), 'truck': (3, 10, 3.5)}
    out = {}
    for k, v in raw.items():
        a = np.array(v, dtype=np.float64)
        a[:2] = a[:2] / step
        out[k] = a
    return out


def _default_anchor_dims():
    d = class_dims()
    small = np.mean(np.stack((d['animal'], d['bicycle'], d['pedestrian'], d['motorcycle'])), axis=0)
    med = d['car']
    large = np.mean(np.stack((d['bus'], d['emergency_vehicle'], d['truck'], d['other_vehicle'])), axis=0)
    return (small, small, med, med, large, large)


cfg = PPConfig()
