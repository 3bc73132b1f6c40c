"""Dataset used by tests/test_gpu_integration.py (a module of its own: DataLoader workers started with "spawn" import it).

``DropInDataset.__getitem__`` follows the reference's PPDataset.__getitem__ (data/dataset.py:88-118) statement by
statement, with the drop-in module functions in place of ``data.pillars`` / ``utils.box_utils``."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


class DropInDataset(torch.utils.data.Dataset):
    P, N, FM = 1200, 16, 40

    def __init__(self, n):
        self.n = n

    def __len__(self):
        return self.n

    def _inputs(self, i):
        import pp_b200
        from pp_b200 import synth
        from oracle import config as ocfg, targets as T
        from helpers import boxes_from_gt
        lidar = np.ascontiguousarray(synth.make_sweep(40 + i)[:6000, :4].astype(np.float64).T).T      # [4,n].T view, dataset.py:88
        mean = torch.from_numpy(synth.make_data_mean(self.P, self.N, seed=2))
        gt = synth.make_gt(40 + i, 10, pp_b200.PPConfig(canvas_width=2 * self.FM, canvas_height=2 * self.FM))
        gt["centers"][:, 1] = 599 - gt["centers"][:, 1]
        boxes = boxes_from_gt(gt, T.Box, ocfg.CLASS_NAMES)
        a_boxes, a_corners, a_centers, _ = T.make_anchor_boxes(self.FM, self.FM)
        return lidar, mean, boxes, a_boxes, a_corners, a_centers

    def __getitem__(self, i):
        from pp_b200 import box_utils, pillars                    # in the worker process: its own CUDA context
        lidar_points, data_mean, boxes, anchor_boxes, anchor_corners, anchor_centers = self._inputs(i)
        pillar = np.zeros((self.P, self.N, 9))
        indices = np.zeros((self.P, 3))
        pillars.create_pillars(lidar_points, pillar, indices, self.N, self.P, .2, .2, -60, -60, -10, 60, 60, 10, np.int32(600))
        pillar = pillar.transpose([2, 0, 1])
        pillar_size = pillar.shape
        pillar = torch.from_numpy(pillar).float()
        pillar = pillar.reshape(-1) - data_mean
        pillar = pillar.reshape(pillar_size)
        indices = torch.from_numpy(indices).long()
        gt_centers, gt_corners = box_utils.boxes_to_image_space(boxes)
        c_target, r_target = box_utils.create_target(anchor_corners, gt_corners, anchor_centers, gt_centers, anchor_boxes, boxes)
        return pillar, indices, torch.from_numpy(c_target).float(), torch.from_numpy(r_target).float()


def oracle_item(ds, i):
    """The same item through the oracle (the reference's compiled pillars.cpp where oracle/_ref holds it)."""
    from oracle import glue, ref, targets as T
    lidar, mean, boxes, a_boxes, a_corners, a_centers = ds._inputs(i)
    m = ref.load()
    pillar, indices = glue.pillarize(lidar, mean, max_pillars=ds.P, max_points=ds.N,
                                     create_pillars=(m.create_pillars if m else None))
    gc, gcor = T.boxes_to_image_space(boxes)
    c, r = T.create_target(a_corners, gcor, a_centers, gc, a_boxes, boxes, make_ious=(m.make_ious if m else None))
    return pillar, indices, torch.from_numpy(c).float(), torch.from_numpy(r).float()
