"""Per-kernel CUDA-event times of target assignment (K3) at the bench shape, both zero-stream variants."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pp_b200
from pp_b200 import _lib, pipeline, synth

L = _lib.load()
path = pipeline.InputPath(device=torch.device("cuda"))
gts = [synth.make_gt(i, 100) for i in range(4)]
batch = path.pack_host_batch([synth.make_sweep(i)[:100] for i in range(4)], gts)
_, g = path.upload(batch)
a = path.ensure_anchors()
out = (torch.empty((4, a.A, 9), device="cuda"), torch.empty((4, a.A, 9), device="cuda"))
ref = None
for mode in (0, 1):
    L.pp_set_option(b"encode_bulk", mode)
    for _ in range(3):
        path.targets(g, batch["gt_offsets"], out=out)
    torch.cuda.synchronize()
    if ref is None:
        ref = (out[0].clone(), out[1].clone())
    else:
        assert torch.equal(ref[0], out[0]) and torch.equal(ref[1], out[1])
    L.pp_profile_enable(1)
    for _ in range(10):
        path.targets(g, batch["gt_offsets"], out=out)
    print("encode_bulk =", mode, {k: round(1e3 * ms / n, 1) for k, (n, ms) in _lib.profile_report().items()})
    L.pp_profile_enable(0)
