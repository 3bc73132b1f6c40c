"""Reduced-size pass over every kernel of libpp_b200.so, the program scripts/sanitize.sh runs under compute-sanitizer:
__graft_entry__.smoke() (K1, K2 dense + fused, K3, backward, loss, aggregation at toy sizes) plus the reference slot
count N = 200 (the specialised tensor-core padding pass), the fused backward, the positives list with a binding
capacity and the loss that consumes it."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import __graft_entry__ as g
import pp_b200
from pp_b200 import box_utils, pipeline, synth
from pp_b200.loss import PPLoss

g.smoke()
cfg = pp_b200.PPConfig(max_pillars=600, max_points_per_pillar=200, fm_height=40, fm_width=40)
P, N = cfg.max_pillars, cfg.max_points_per_pillar
mean = synth.make_data_mean(P, N, seed=3)
for fused in (True, False):
    path = pipeline.InputPath(cfg, data_mean=mean, pfn_params=synth.make_pfn_params(2, flip_gamma=True), training=True, fused=fused)
    sweeps = [synth.make_sweep(s)[:20000] for s in (3, 4)]
    gcfg = pp_b200.PPConfig(canvas_width=80, canvas_height=80)
    gts = [synth.make_gt(s, 12, gcfg) for s in (3, 4)]
    for gt in gts:
        gt["centers"][:, 1] = 599 - gt["centers"][:, 1]
    batch = path.pack_host_batch(sweeps, gts)
    for _ in range(2):
        h = path.step_host_async(batch)
        h.counters()
    if fused:
        pts, gt_dev = path.upload(batch)
        canvas, inds, npil = path.pillarize_encode_train(pts, batch["offsets"])
        canvas.backward(torch.randn_like(canvas))
        a = path.ensure_anchors()
        pos, _, _, _ = box_utils.assign_targets(a, gt_dev["corners"], gt_dev["centers"], gt_dev["wlh"], gt_dev["yaw"], gt_dev["cls"],
                                                batch["gt_offsets"], as_list=True, capacity=6)
        cls = torch.randn((2, 54, 40, 40), device="cuda") - 3.0
        reg = torch.randn((2, 48, 40, 40), device="cuda")
        out = PPLoss(0.4, 1.0, 250.0, 2, torch.device("cuda"))(cls.requires_grad_(True), reg.requires_grad_(True) * 1.0, pos)
        out[4].backward()
        try:
            path.check_status()
        except pp_b200._lib.PPError as e:
            print("expected:", str(e)[:80])
torch.cuda.synchronize()
print("sanitize target ok")
