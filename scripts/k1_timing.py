"""(needs a debug build: PP_DEBUG=1 python 3d-object-detection_b200/build.py)  Stage boundaries of block 0 of k_pillarize: python scripts/k1_timing.py"""
import ctypes
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import pp_b200
from pp_b200 import _lib, pipeline, synth

L = _lib.load()
B, P, N = 4, 24000, 200
path = pipeline.InputPath(data_mean=synth.make_data_mean(P, N, dense=True), pfn_params=synth.make_pfn_params(0), fused=True)
sweeps = [synth.make_sweep(i) for i in range(B)]
offs = np.cumsum([0] + [len(s) for s in sweeps]).tolist()
pts = torch.from_numpy(np.concatenate(sweeps)).cuda()
for _ in range(3):
    path.pillarize_encode(pts, offs)
torch.cuda.synchronize()
names = ["bin", "barrier", "tilecount", "assign", "barrier", "scatter", "barrier", "rank", "barrier", "mean", "barrier", "feat"]
acc = np.zeros(len(names))
accm = np.zeros(len(names))
R = 5
dbg = int(sys.argv[1]) if len(sys.argv) > 1 else 0
for _ in range(R):
    L.pp_debug_set(b"pfn_tc_timing", 1)
    L.pp_debug_set(b"pfn_tc_debug", dbg)
    path.pillarize_encode(pts, offs, stages=1)     # pillarize stage only: the PFN kernels share the timing buffer
    torch.cuda.synchronize()
    buf = (ctypes.c_int64 * 128)()
    L.pp_debug_tc_timing(buf)
    L.pp_debug_set(b"pfn_tc_timing", 0)
    L.pp_debug_set(b"pfn_tc_debug", 0)
    t = np.array([buf[i] for i in range(len(names) + 1)], dtype=np.float64)
    tm = np.array([buf[32 + i] for i in range(len(names) + 1)], dtype=np.float64)
    acc += np.diff(t) / 1e3
    accm += (tm[1:] - t[:-1]) / 1e3
print("%-10s %10s %22s" % ("stage", "block 0", "block 0 start -> slowest block end"))
for n, v, w in zip(names, acc / R, accm / R):
    print("%-10s %7.2f us %7.2f us" % (n, v, w))
print("total %.2f us" % (acc.sum() / R))
