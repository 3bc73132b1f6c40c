"""Host-side cost of InputPath.step_host_async (Python + ctypes + torch launches) per step, with cProfile."""
import cProfile, pstats, io, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pp_b200
from pp_b200 import pipeline, synth

cfg = pp_b200.PPConfig()
P, N = cfg.max_pillars, cfg.max_points_per_pillar
path = pipeline.InputPath(cfg, device=torch.device("cuda"), data_mean=synth.make_data_mean(P, N, seed=0, dense=True),
                          pfn_params=synth.make_pfn_params(0), training=True, fused=True)
sweeps = [synth.make_sweep(i) for i in range(4)]
gts = [synth.make_gt(i, 100) for i in range(4)]
batch = path.pack_host_batch(sweeps, gts)
A = path.ensure_anchors().A
def alloc():
    return {"pillars": (None, torch.empty((4, P, 3), dtype=torch.int64, device="cuda"), torch.empty(4, dtype=torch.int32, device="cuda")),
            "canvas": torch.empty((4, 64, 600, 600), device="cuda"),
            "targets": (torch.empty((4, A, 9), device="cuda"), torch.empty((4, A, 9), device="cuda"))}
outs = [alloc(), alloc()]
pend = []
def loop(n):
    for i in range(n):
        pend.append(path.step_host_async(batch, out=outs[i & 1]))
        if len(pend) > 1:
            pend.pop(0).counters()
    while pend:
        pend.pop(0).counters()
loop(10)
torch.cuda.synchronize()
# pure issue cost: no waiting on results (steps pile up on the GPU)
t0 = time.perf_counter()
hs = [path.step_host_async(batch, out=outs[i & 1]) for i in range(6)]
t1 = time.perf_counter()
torch.cuda.synchronize()
print("host issue time per step (no waits): %.1f us" % ((t1 - t0) / 6 * 1e6))
pr = cProfile.Profile()
pr.enable()
loop(200)
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(18)
print(s.getvalue()[:3500])
