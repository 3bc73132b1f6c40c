"""Per-kernel CUDA-event times of the fused input path at the bench shape (product build): python scripts/kernel_times.py [B]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import pp_b200
from pp_b200 import _lib, pipeline, synth

L = _lib.load()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
P, N = 24000, 200
path = pipeline.InputPath(data_mean=synth.make_data_mean(P, N, dense=True), pfn_params=synth.make_pfn_params(0), fused=True)
sweeps = [synth.make_sweep(i) for i in range(B)]
offs = np.cumsum([0] + [len(s) for s in sweeps]).tolist()
pts = torch.from_numpy(np.concatenate(sweeps)).cuda()
for _ in range(3):
    path.pillarize_encode(pts, offs)
torch.cuda.synchronize()
L.pp_profile_enable(1)
for _ in range(8):
    path.pillarize_encode(pts, offs)
tot = 0.0
for k, (n, ms) in _lib.profile_report().items():
    print("%-18s %8.2f us" % (k, 1e3 * ms / n))
    tot += 1e3 * ms / n * (n / 8)
print("sum %.1f us per batch" % tot)
