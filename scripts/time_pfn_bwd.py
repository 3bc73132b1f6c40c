"""Times pp_pfn_backward (training mode) at the reference's shape, B sweeps of [9, 24000, 200]."""
import argparse, json, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pp_b200
from pp_b200 import _lib, model as pm


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--iters", type=int, default=5)
    a = ap.parse_args()
    B, P, N = a.batch, 24000, 200
    x = torch.randn((B, 9, P, N), device="cuda")
    net = pm.PPFeatureNet(9, 64).cuda().train()
    g = torch.randn((B, 64, P), device="cuda")
    L = _lib.load()
    for _ in range(2):
        net.zero_grad(); net(x).backward(g)
    torch.cuda.synchronize()
    L.pp_profile_enable(1)
    for _ in range(a.iters):
        net.zero_grad(); net(x).backward(g)
    rep = {k: round(ms * 1e3 / n, 1) for k, (n, ms) in _lib.profile_report().items()}
    L.pp_profile_enable(0)
    # the fused (x-free) path: forward pp_input_path + backward pp_input_path_backward on synthetic sweeps
    from pp_b200 import pipeline, synth
    path = pipeline.InputPath(device=torch.device("cuda"), data_mean=synth.make_data_mean(P, N, seed=0, dense=True),
                              pfn_params=synth.make_pfn_params(0), training=True, fused=True)
    sweeps = [synth.make_sweep(i) for i in range(B)]
    pts = torch.tensor(np.concatenate(sweeps), device="cuda")
    offs = [0] + list(np.cumsum([len(s_) for s_ in sweeps]))
    gc = torch.randn((B, 64, 600, 600), device="cuda")
    for _ in range(2):
        path.net.zero_grad(); path.pillarize_encode_train(pts, offs)[0].backward(gc)
    torch.cuda.synchronize()
    L.pp_profile_enable(1)
    for _ in range(a.iters):
        path.net.zero_grad(); path.pillarize_encode_train(pts, offs)[0].backward(gc)
    rep2 = {k: round(ms * 1e3 / n, 1) for k, (n, ms) in _lib.profile_report().items() if k.startswith("k_pfn_bwd")}
    L.pp_profile_enable(0)
    print(json.dumps({"batch": B, "kernels_us": rep, "fused_backward_kernels_us": rep2}))


if __name__ == "__main__":
    main()
