"""Per-kernel CUDA-event times of the signature-preserving dense path (product build)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import pp_b200
from pp_b200 import _lib, pipeline, synth

L = _lib.load()
B, P, N = 4, 24000, 200
path = pipeline.InputPath(data_mean=synth.make_data_mean(P, N, dense=True), pfn_params=synth.make_pfn_params(0), fused=False)
sweeps = [synth.make_sweep(i) for i in range(B)]
offs = np.cumsum([0] + [len(s) for s in sweeps]).tolist()
x, inds, npil = path.pillarize(torch.from_numpy(np.concatenate(sweeps)).cuda(), offs)
canvas = torch.empty((B, 64, 600, 600), device="cuda")
for _ in range(3):
    path.encode(x, inds, out=canvas)
torch.cuda.synchronize()
L.pp_profile_enable(1)
for _ in range(8):
    path.encode(x, inds, out=canvas)
for k, (n, ms) in _lib.profile_report().items():
    print("%-18s %8.2f us" % (k, 1e3 * ms / n))
