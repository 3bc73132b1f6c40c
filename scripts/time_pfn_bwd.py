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
    print(json.dumps({"batch": B, "kernels_us": rep}))


if __name__ == "__main__":
    main()
