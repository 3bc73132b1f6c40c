"""(needs a debug build: PP_DEBUG=1 python 3d-object-detection_b200/build.py)  Run the PFN (+scatter) stage a few times (for ncu captures of k_pfn_stats_tc / k_canvas)."""
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
path.net.train(True)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
    path.encode(x, inds, out=canvas)
torch.cuda.synchronize()
print("ok")
if len(sys.argv) > 2:
    import ctypes
    L.pp_debug_set(b"pfn_tc_timing", 1)
    if len(sys.argv) > 3: L.pp_debug_set(b"pfn_tc_debug", int(sys.argv[3]))
    path.encode(x, inds, out=canvas)
    buf = (ctypes.c_int64 * 128)()
    L.pp_debug_tc_timing(buf)
    L.pp_debug_set(b"pfn_tc_timing", 0)
    for w in range(26):
        v = [buf[w * 4 + k] for k in range(4)]
        role = ("epi e%d j%d q%d (acc_full,busy)" % (w >> 3, (w >> 2) & 1, w & 3) if w < 16 else "conv(raw_full,b_empty)" if w < 24
                else "producer(raw_empty,-)" if w == 24 else "mma(b_full,acc_empty,issue)")
        print("warp %2d %-30s wait0 %8d wait1 %8d x2 %8d total %8d" % (w, role, v[0], v[1], v[2], v[3]))
