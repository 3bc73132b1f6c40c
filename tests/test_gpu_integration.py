"""The drop-ins at the reference's own call sites.

(a) The reference's PPModel (model/model.py:162-180, the reference's own module loaded by oracle/refmodel.py: sources
    in the build container, byte-compiled copy on the GPU box) with ``feature_net`` / ``scatter`` swapped for
    pp_b200.model.PPFeatureNet / PPScatter, one state_dict loaded into both, (cls, reg) head outputs compared on the
    GPU with TF32 off.
(b) A DataLoader(num_workers=2, multiprocessing_context="spawn") whose __getitem__ does what the reference's
    PPDataset.__getitem__ does (data/dataset.py:88-118) with the drop-in module functions:
    pillars.create_pillars -> transpose / float / - data_mean -> boxes_to_image_space -> create_target.  The reference
    forks its workers (train.py:120-121); a CUDA context does not survive fork, so the drop-in needs "spawn"
     -- this test is that configuration, against the oracle computed in the parent."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_reference_ppmodel_with_dropin_feature_net_and_scatter():
    from oracle import refmodel
    loaded = refmodel.load()
    if loaded is None:
        pytest.skip("reference modules not available (oracle/_ref/refpy not built)")
    mod, rcfg, kind = loaded
    import pp_b200
    from pp_b200 import model as M, synth
    from oracle import glue
    tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        dev = torch.device("cuda")
        torch.manual_seed(3)
        ref = mod.PPModel(9, 64, 54, 48, dev).to(dev)
        ours = mod.PPModel(9, 64, 54, 48, dev).to(dev)
        ours.load_state_dict(ref.state_dict())
        # the swap a user of the reference makes: same attribute names, same state_dict keys
        fnet, scat = M.PPFeatureNet(9, 64).to(dev), M.PPScatter(dev)
        missing = fnet.load_state_dict(ref.feature_net.state_dict(), strict=True)
        assert not missing.missing_keys and not missing.unexpected_keys
        ours.feature_net, ours.scatter = fnet, scat
        assert set(ours.state_dict().keys()) == set(ref.state_dict().keys())
        P, N = 2000, 32                                  # network input of two small sweeps, reference layout
        mean = synth.make_data_mean(P, N, seed=4, dense=True)
        xs, iis = zip(*[glue.pillarize(synth.make_sweep(s)[:9000, :4].astype(np.float64), torch.from_numpy(mean),
                                       max_pillars=P, max_points=N) for s in (5, 6)])
        x, inds = torch.stack(xs).to(dev), torch.stack(iis).to(dev)
        for training in (True, False):
            ref.train(training); ours.train(training)
            with torch.no_grad():
                c0, r0 = ref(x, inds)
                c1, r1 = ours(x, inds)
            assert c0.shape == c1.shape == (2, 54, 300, 300) and r0.shape == r1.shape == (2, 48, 300, 300)
            for a, b in ((c0, c1), (r0, r1)):
                scale = float(a.abs().max())
                d = float((a - b).abs().max())
                print("%s training=%s: max |diff| %.3g of max |out| %.3g" % (kind, training, d, scale))
                assert d <= 2e-4 * scale + 1e-5          # 16 conv + BatchNorm layers of fp32 behind a 1e-5-relative canvas
        # training mode updated the running statistics of both feature nets identically
        assert torch.allclose(ref.feature_net.bn1.running_mean, fnet.bn1.running_mean, rtol=1e-5, atol=1e-6)
        assert torch.allclose(ref.feature_net.bn1.running_var, fnet.bn1.running_var, rtol=1e-5, atol=1e-6)
        assert int(ref.feature_net.bn1.num_batches_tracked) == int(fnet.bn1.num_batches_tracked) == 1
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32


def test_spawn_dataloader_getitem_through_the_dropin_functions():
    from integration_dataset import DropInDataset, oracle_item
    ds = DropInDataset(n=4)
    dl = torch.utils.data.DataLoader(ds, batch_size=2, num_workers=2, multiprocessing_context="spawn", shuffle=False)
    got = list(dl)
    assert len(got) == 2
    pillar = torch.cat([g[0] for g in got]); inds = torch.cat([g[1] for g in got])
    c_t = torch.cat([g[2] for g in got]); r_t = torch.cat([g[3] for g in got])
    assert pillar.shape == (4, 9, ds.P, ds.N) and pillar.dtype == torch.float32 and inds.dtype == torch.int64
    for i in range(4):
        p0, i0, c0, r0 = oracle_item(ds, i)
        assert torch.equal(pillar[i], p0) and torch.equal(inds[i], i0)          # bit-exact network input
        assert torch.equal(c_t[i], c0)                                          # labels bit-exact
        assert torch.equal(r_t[i][:, 0], r0[:, 0]) and torch.equal(r_t[i][:, 8], r0[:, 8])
        a, b = r_t[i][:, 1:8].double(), r0[:, 1:8].double()
        assert bool(((a - b).abs() <= 1e-5 * torch.maximum(a.abs(), b.abs()) + 1e-7).all())
