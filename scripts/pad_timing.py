"""(needs a debug build: PP_DEBUG=1 python 3d-object-detection_b200/build.py)  Role timing of CTA 0 of k_pfn_pad_tc (development): python scripts/pad_timing.py"""
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
dbg0 = int(sys.argv[1]) if len(sys.argv) > 1 else 0
L.pp_debug_set(b"pfn_tc_debug", dbg0)
L.pp_debug_set(b"pfn_tc_timing", 1)
path.pillarize_encode(pts, offs)
buf = (ctypes.c_int64 * 128)()
L.pp_debug_tc_timing(buf)
L.pp_debug_set(b"pfn_tc_timing", 0)
L.pp_debug_set(b"pfn_tc_debug", 0)
print("role timing with dbg =", dbg0)
for w in (0, 4, 8, 12, 16, 17):
    v = [buf[w * 4 + k] for k in range(4)]
    role = ("epi j%d q%d (wait acc_full, busy, wait pbar)" % (w >> 2, w & 3) if w < 16 else
            "producer (wait empty)" if w == 16 else "mma (wait full, wait acc_empty, issue)")
    print("warp %2d %-44s %9d %9d %9d total %9d" % (w, role, v[0], v[1], v[2], v[3]))
L.pp_profile_enable(1)
for dbg in (0, 2, 4, 6, 22):
    L.pp_debug_set(b"pfn_tc_debug", dbg)      # bit 1: no MMAs, bit 2: no epilogue loads / arithmetic
    for _ in range(5):
        path.pillarize_encode(pts, offs)
    rep = _lib.profile_report()
    if dbg == 0:
        for k, (n, ms) in rep.items():
            print("%-18s %8.2f us" % (k, 1e3 * ms / n))
    print("dbg=%d k_pfn_pad_tc %8.2f us  k_pfn_real %8.2f us" % (dbg, 1e3 * rep["k_pfn_pad_tc"][1] / rep["k_pfn_pad_tc"][0],
                                                                 1e3 * rep["k_pfn_real"][1] / rep["k_pfn_real"][0]))
L.pp_debug_set(b"pfn_tc_debug", 0)
