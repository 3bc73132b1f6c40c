"""(needs a debug build: PP_DEBUG=1 python 3d-object-detection_b200/build.py)  Time the PFN (+scatter) stage alone, train vs eval mode, tensor-core vs CUDA-core kernel."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import pp_b200
from pp_b200 import _lib, pipeline, synth
L = _lib.load()
B, P, N = 4, 24000, 200
path = pipeline.InputPath(data_mean=synth.make_data_mean(P, N, dense=True), pfn_params=synth.make_pfn_params(0))
sweeps = [synth.make_sweep(i) for i in range(B)]
offs = np.cumsum([0] + [len(s) for s in sweeps]).tolist()
x, inds, npil = path.pillarize(torch.from_numpy(np.concatenate(sweeps)).cuda(), offs)
canvas = torch.empty((B, 64, 600, 600), device="cuda")
def run(train, tc, dense=True):
    L.pp_set_option(b"pfn_tensor_cores", tc)
    path.net.train(train)
    xx = x if dense else xz
    for _ in range(3): path.encode(xx, inds, out=canvas)
    L.pp_profile_enable(1)
    for _ in range(10): path.encode(xx, inds, out=canvas)
    rep = _lib.profile_report(); L.pp_profile_enable(0)
    return {k: round(v[1] / v[0] * 1e3, 1) for k, v in rep.items()}
path0 = pipeline.InputPath(pfn_params=synth.make_pfn_params(0))
xz, _, _ = path0.pillarize(torch.from_numpy(np.concatenate(sweeps)).cuda(), offs)
for train in (True, False):
    for tc in (1, 0):
        print("train=%s tc=%d dense-mean:" % (train, tc), run(train, tc))
print("train=True tc=0 no data_mean (zero slots skipped):", run(True, 0, dense=False))
print("train=True tc=1 no data_mean:", run(True, 1, dense=False))

for dbg in (1, 2, 4, 3, 5, 6, 7):
    L.pp_debug_set(b"pfn_tc_debug", dbg)
    print("dbg=%d (1=no convert, 2=no mma, 4=no epilogue reads) train:" % dbg, run(True, 1)["k_pfn_stats_tc"], "eval:", run(False, 1)["k_pfn_stats_tc"])
L.pp_debug_set(b"pfn_tc_debug", 0)
