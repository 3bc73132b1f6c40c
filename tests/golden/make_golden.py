"""Generate the committed golden fixtures from the REFERENCE itself.  Run in the build container
only (needs /root/reference); the fixtures travel, the reference does not.

    python tests/golden/make_golden.py

* config.json          values the reference's config.py evaluates to (anchor dims, grid, caps)
* pfn_small.npz        reference model/model.py PPFeatureNet (fp64, CPU) in train and eval mode +
                       PPScatter non-zeros, imported with a 6-line ``easydict`` stand-in
* pillars_small.npz    reference data/pillars.cpp (compiled against oracle/boost_shim, see
                       oracle/Makefile) create_pillars / make_ious on small seeded inputs
"""
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _easydict_standin():
    mod = types.ModuleType("easydict")

    class EasyDict(dict):
        def __getattr__(self, k):
            try:
                return self[k]
            except KeyError as e:
                raise AttributeError(k) from e

        def __setattr__(self, k, v):
            self[k] = v
    mod.EasyDict = EasyDict
    sys.modules["easydict"] = mod


def main():
    assert os.path.isdir(REF), "reference checkout not present"
    _easydict_standin()
    sys.path.insert(0, REF)
    import config as refcfg                      # the reference's own config.py
    from model.model import PPFeatureNet, PPScatter   # the reference's own modules

    c = refcfg.cfg
    cfg_out = {
        "X_MIN": c.DATA.X_MIN, "Y_MIN": c.DATA.Y_MIN, "Z_MIN": c.DATA.Z_MIN,
        "X_MAX": c.DATA.X_MAX, "Y_MAX": c.DATA.Y_MAX, "Z_MAX": c.DATA.Z_MAX,
        "X_STEP": c.DATA.X_STEP, "Y_STEP": c.DATA.Y_STEP, "FM_SCALE": c.DATA.FM_SCALE,
        "FM_HEIGHT": int(c.DATA.FM_HEIGHT), "FM_WIDTH": int(c.DATA.FM_WIDTH),
        "CANVAS_HEIGHT": int(c.DATA.CANVAS_HEIGHT), "CANVAS_WIDTH": int(c.DATA.CANVAS_WIDTH),
        "ANCHOR_DIMS": [list(map(float, d)) for d in c.DATA.ANCHOR_DIMS],
        "ANCHOR_YAWS": list(c.DATA.ANCHOR_YAWS), "ANCHOR_ZS": list(c.DATA.ANCHOR_ZS),
        "MAX_POINTS_PER_PILLAR": c.DATA.MAX_POINTS_PER_PILLAR, "MAX_PILLARS": c.DATA.MAX_PILLARS,
        "REG_DIMS": c.DATA.REG_DIMS, "IOU_POS_THRESH": c.DATA.IOU_POS_THRESH,
        "NUM_CLASSES": c.DATA.NUM_CLASSES, "NAME_TO_IND": dict(c.DATA.NAME_TO_IND),
        "FEATURE_NET_IN": c.NET.FEATURE_NET_IN, "FEATURE_NET_OUT": c.NET.FEATURE_NET_OUT,
    }
    with open(os.path.join(HERE, "config.json"), "w") as f:
        json.dump(cfg_out, f, indent=1, sort_keys=True)

    # ---- PFN / scatter from the reference modules, fp64 on CPU --------------------------------
    torch.manual_seed(1234)
    B, D, P, N, C = 2, 9, 48, 12, 64
    rng = np.random.default_rng(7)
    x = np.zeros((B, D, P, N), np.float32)
    occ = rng.integers(0, N + 1, (B, P))
    occ[:, 40:] = 0                                        # trailing empty pillars
    for b in range(B):
        for p in range(P):
            k = occ[b, p]
            x[b, :, p, :k] = rng.normal(0, 3, (D, k)).astype(np.float32)
    x -= rng.normal(0, 0.05, (1, D, P, N)).astype(np.float32)   # a per-slot "data_mean"
    inds = np.zeros((B, P, 3), np.int64)
    cells = rng.choice(600 * 600, (B, 40), replace=False)
    inds[:, :40, 0] = 1
    inds[:, :40, 1] = cells % 600
    inds[:, :40, 2] = cells // 600
    out = {}
    for tag, flip in (("pos", False), ("mixed", True)):
        net = PPFeatureNet(D, C).double()
        with torch.no_grad():
            if flip:
                g = torch.from_numpy(rng.uniform(0.5, 1.5, C) * rng.choice([-1.0, 1.0], C))
                net.bn1.weight.copy_(g)
                net.bn1.bias.copy_(torch.from_numpy(rng.normal(0, 0.2, C)))
                net.bn1.running_mean.copy_(torch.from_numpy(rng.normal(1, 0.5, C)))
                net.bn1.running_var.copy_(torch.from_numpy(rng.uniform(0.5, 3, C)))
        # parameters are stored as float32 values so that a float32 module holds exactly them
        with torch.no_grad():
            for prm in list(net.parameters()) + [net.bn1.running_mean, net.bn1.running_var]:
                prm.copy_(prm.float().double())
        sd0 = {k: v.clone().numpy() for k, v in net.state_dict().items()}
        xt = torch.from_numpy(x).double()
        net.eval()
        with torch.no_grad():
            y_eval = net(xt).numpy()
        net.train()
        with torch.no_grad():
            y_train = net(xt).numpy()
        sd1 = {k: v.clone().numpy() for k, v in net.state_dict().items()}
        for k, v in sd0.items():
            out["%s/sd0/%s" % (tag, k)] = v
        out["%s/y_eval" % tag] = y_eval
        out["%s/y_train" % tag] = y_train
        out["%s/rm1" % tag] = sd1["bn1.running_mean"]
        out["%s/rv1" % tag] = sd1["bn1.running_var"]
        out["%s/nbt1" % tag] = sd1["bn1.num_batches_tracked"]
        if tag == "pos":
            sc = PPScatter(torch.device("cpu"))
            canvas = sc(torch.from_numpy(y_train).float(), torch.from_numpy(inds)).numpy()   # the reference canvas is float32
            nz = np.nonzero(canvas)
            out["scatter/shape"] = np.array(canvas.shape)
            out["scatter/nz_index"] = np.stack(nz, 1).astype(np.int32)
            out["scatter/nz_value"] = canvas[nz]
    out["x"] = x
    out["inds"] = inds
    np.savez_compressed(os.path.join(HERE, "pfn_small.npz"), **out)

    # ---- create_pillars / make_ious from the reference's pillars.cpp --------------------------
    from oracle import ref
    ref.build()
    m = ref.load()
    assert m is not None
    from helpers import GRID, cloud_boundaries, cloud_dense_cells, cloud_random, run_create_pillars
    g = {}
    for name, pts, P_, N_ in (("boundaries", cloud_boundaries(), 16, 4),
                              ("random", cloud_random(21, n=1500, cols=4), 600, 6),
                              ("dense", cloud_dense_cells(22, n=1200, ncells=6), 8, 40),
                              ("capP", cloud_random(23, n=1500, spread=59.0, cols=4), 100, 5)):
        t, ind = run_create_pillars(m.create_pillars, pts, P_, N_)
        nzp = np.nonzero(ind[:, 0])[0]
        g[name + "/points"] = pts
        g[name + "/PN"] = np.array([P_, N_])
        g[name + "/indices"] = ind
        g[name + "/tensor_nz_index"] = np.stack(np.nonzero(t), 1).astype(np.int32)
        g[name + "/tensor_nz_value"] = t[np.nonzero(t)]
        g[name + "/n_pillars"] = np.array(len(nzp))
    from test_oracle_iou import cw, rect
    rng = np.random.default_rng(31)
    A, G = 300, 6
    ac = np.stack([rng.uniform(0, 40, A), rng.uniform(0, 40, A), np.zeros(A)], 1)
    gc = np.stack([rng.uniform(0, 40, G), rng.uniform(0, 40, G), np.zeros(G)], 1)
    a_cor = np.stack([rect(ac[i, 0], ac[i, 1], 8, 4, rng.uniform(-3, 3)) for i in range(A)])
    g_cor = np.stack([cw(rect(gc[j, 0], gc[j, 1], 9, 4, rng.uniform(-3, 3))) for j in range(G)])
    ious = np.zeros((A, G))
    m.make_ious(a_cor, g_cor, ac, gc, ious)
    g["iou/a_corners"], g["iou/g_corners"], g["iou/a_centers"], g["iou/g_centers"] = a_cor, g_cor, ac, gc
    g["iou/ious"] = ious
    np.savez_compressed(os.path.join(HERE, "pillars_small.npz"), **g)
    print("fixtures written to", HERE)


if __name__ == "__main__":
    main()
