"""Run the fused input path (pp_input_path) a few times at the bench shape, for ncu captures of
k_pfn_pad_tc / k_pfn_real / k_canvas / the K1 stages.  Usage: python scripts/run_fused_once.py [steps] [B]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import pp_b200
from pp_b200 import pipeline, synth

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4
P, N = 24000, 200
path = pipeline.InputPath(data_mean=synth.make_data_mean(P, N, dense=True), pfn_params=synth.make_pfn_params(0), fused=True)
sweeps = [synth.make_sweep(i) for i in range(B)]
offs = np.cumsum([0] + [len(s) for s in sweeps]).tolist()
pts = torch.from_numpy(np.concatenate(sweeps)).cuda()
out = {"canvas": torch.empty((B, 64, 600, 600), device="cuda"),
       "pillars": (None, torch.empty((B, P, 3), dtype=torch.int64, device="cuda"), torch.empty(B, dtype=torch.int32, device="cuda"))}
path.net.train(True)
for _ in range(steps):
    path.pillarize_encode(pts, offs, out=out)
torch.cuda.synchronize()
print("ok", out["pillars"][2].tolist())
