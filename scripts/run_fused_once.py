"""Run the fused K1+K2 path a few times (for ncu captures)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
from pp_b200 import _lib, pipeline, synth
P, N = 24000, 200
path = pipeline.InputPath(data_mean=synth.make_data_mean(P, N, dense=True), pfn_params=synth.make_pfn_params(0), training=True)
sweeps = [synth.make_sweep(s) for s in range(4)]
offs = np.cumsum([0] + [len(s) for s in sweeps]).tolist()
pts = torch.from_numpy(np.concatenate(sweeps)).cuda()
out = {"canvas": torch.empty((4, 64, 600, 600), device="cuda")}
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4): path.pillarize_encode(pts, offs, out=out)
torch.cuda.synchronize(); print("ok")
