"""Probe: throughput of the fused step when two independent steps are in flight on two streams."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
from pp_b200 import _lib, pipeline, synth
P, N = 24000, 200
mean = synth.make_data_mean(P, N, dense=True)
prm = synth.make_pfn_params(0)
B = 4
sweeps = [synth.make_sweep(s) for s in range(B)]
gts = [synth.make_gt(s, 100) for s in range(B)]
lanes = []
for i in range(2):
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        path = pipeline.InputPath(data_mean=mean, pfn_params=prm, training=True, fused=True)
        batch = path.pack_host_batch(sweeps, gts)
        d_pts, gt_dev = path.upload(batch)
        A = path.ensure_anchors().A
        out = {"pillars": (None, torch.empty((B, P, 3), dtype=torch.int64, device="cuda"), torch.empty(B, dtype=torch.int32, device="cuda")),
               "canvas": torch.empty((B, 64, 600, 600), device="cuda"),
               "targets": (torch.empty((B, A, 9), device="cuda"), torch.empty((B, A, 9), device="cuda"))}
    lanes.append((st, path, batch, d_pts, gt_dev, out))
torch.cuda.synchronize()
def run(nl, steps):
    for i in range(steps):
        st, path, batch, d_pts, gt_dev, out = lanes[i % nl]
        with torch.cuda.stream(st):
            path.step_device(d_pts, batch["offsets"], gt_dev, batch["gt_offsets"], out=out)
for nl in (1, 2, 1, 2):
    run(nl, 6); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(torch.cuda.default_stream())
    import time; t0 = time.perf_counter()
    run(nl, 40)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("lanes %d: %.1f us per step (host wall), %.0f sweeps/s" % (nl, dt / 40 * 1e6, B * 40 / dt))
# host-only enqueue cost: enqueue while the GPU is busy with a long kernel backlog
torch.cuda.synchronize()
big = torch.empty(1 << 28, device="cuda")
for _ in range(20): big.normal_()           # ~ tens of ms of queued GPU work
t0 = time.perf_counter()
run(1, 20)
dt = time.perf_counter() - t0
print("host enqueue time per step: %.1f us" % (dt / 20 * 1e6))
torch.cuda.synchronize()
