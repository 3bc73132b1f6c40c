"""(needs a debug build: PP_DEBUG=1 python 3d-object-detection_b200/build.py)  Print the fused-vs-dense canvas differences and time the fused step (development aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import pp_b200
from pp_b200 import _lib, pipeline, synth
L = _lib.load()
P, N = 24000, 200
for dense_mean in (True, False):
    for training in (True, False):
        mean = synth.make_data_mean(P, N, dense=dense_mean)
        prm = synth.make_pfn_params(1, flip_gamma=True)
        mk = lambda: pipeline.InputPath(data_mean=mean, pfn_params=prm, training=training)
        pa, pb = mk(), mk()
        sweeps = [synth.make_sweep(40 + s) for s in range(4)]
        offs = np.cumsum([0] + [len(s) for s in sweeps]).tolist()
        pts = torch.from_numpy(np.concatenate(sweeps)).cuda()
        x, inds, npil = pa.pillarize(pts, offs)
        want = pa.encode(x, inds)
        canvas, inds2, npil2 = pb.pillarize_encode(pts, offs)
        d = (canvas.double() - want.double()).abs()
        nz = want != 0
        print("mean=%s train=%s max|diff| %.3g  max|canvas| %.3g  mean|diff| on support %.3g  rm diff %.3g rv diff %.3g" % (
            dense_mean, training, d.max().item(), want.abs().max().item(), d[nz].mean().item(),
            (pa.net.bn1.running_mean - pb.net.bn1.running_mean).abs().max().item(),
            (pa.net.bn1.running_var - pb.net.bn1.running_var).abs().max().item()))
# timing
mean = synth.make_data_mean(P, N, dense=True)
path = pipeline.InputPath(data_mean=mean, pfn_params=synth.make_pfn_params(0), training=True)
sweeps = [synth.make_sweep(s) for s in range(4)]
offs = np.cumsum([0] + [len(s) for s in sweeps]).tolist()
pts = torch.from_numpy(np.concatenate(sweeps)).cuda()
out = {"canvas": torch.empty((4, 64, 600, 600), device="cuda")}
for _ in range(3): path.pillarize_encode(pts, offs, out=out)
L.pp_profile_enable(1)
for _ in range(10): path.pillarize_encode(pts, offs, out=out)
rep = _lib.profile_report(); L.pp_profile_enable(0)
tot = 0
for k, v in rep.items():
    print("%-20s %7.1f us" % (k, v[1] / v[0] * 1e3)); tot += v[1] / v[0] * 1e3
print("sum %.1f us" % tot)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): path.pillarize_encode(pts, offs, out=out)
e1.record(); torch.cuda.synchronize()
print("fused K1+K2: %.1f us per batch-4" % (e0.elapsed_time(e1) / 20 * 1e3))
for dbg in (0, 16, 7, 23):
    L.pp_debug_set(b"pfn_tc_debug", dbg)
    for _ in range(2): path.pillarize_encode(pts, offs, out=out)
    L.pp_profile_enable(1)
    for _ in range(5): path.pillarize_encode(pts, offs, out=out)
    rep = _lib.profile_report(); L.pp_profile_enable(0)
    print("dbg=%d k_pfn_pad_tc %.1f us avg over %d launches" % (dbg, rep["k_pfn_pad_tc"][1] / rep["k_pfn_pad_tc"][0] * 1e3, rep["k_pfn_pad_tc"][0]))
L.pp_debug_set(b"pfn_tc_debug", 0)
import ctypes
L.pp_debug_set(b"pfn_tc_timing", 1)
path.pillarize_encode(pts, offs, out=out)
buf = (ctypes.c_int64 * 128)()
L.pp_debug_tc_timing(buf)
L.pp_debug_set(b"pfn_tc_timing", 0)
for w in (0, 4, 8, 12, 16, 20, 24, 25):
    v = [buf[w * 4 + k] for k in range(4)]
    role = ("epi e%d j%d q%d (acc_full,busy)" % (w >> 3, (w >> 2) & 1, w & 3) if w < 16 else "conv(raw_full,b_empty)" if w < 24
            else "producer(raw_empty,-)" if w == 24 else "mma(b_full,acc_empty,issue)")
    print("warp %2d %-30s wait0 %8d wait1 %8d x2 %8d total %8d" % (w, role, v[0], v[1], v[2], v[3]))
