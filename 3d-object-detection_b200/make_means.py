"""Training-set mean of the network input tensor over this pillarizer -- the equivalent of the reference's
make_means.py:28-37 (``pillar_means.pkl``).

``data_mean`` is indexed by pillar SLOT (flat index ``d*P*N + p*N + n``, data/dataset.py:99-105) and the slot a
pillar lands in depends on the pillarizer's pillar order.  The reference emits pillars in the iteration order of a
Boost ``unordered_map`` (data/pillars.cpp:332, implementation-defined); this library emits them in first-touch order
(DESIGN.md section 2).  A ``pillar_means.pkl`` produced by the reference pipeline is therefore NOT numerically
interchangeable with this path: regenerate it with this function (same running-mean recurrence, same float32
arithmetic as the reference script), and re-validate or fine-tune a checkpoint that was trained against the
reference's means.
"""
import torch

from . import pipeline
from .config import PPConfig


def make_means(sample_batches, cfg=None, device=None):
    """``sample_batches``: iterable of lists of float32 ``[n_i, >=4]`` point clouds (one list = one mini-batch, the
    reference uses 5 samples per batch).  Returns the float32 CPU tensor ``[9*P*N]`` that ``InputPath(data_mean=...)``
    and the reference's ``PPDataset(data_mean=...)`` take:  means = means*(i/(i+1)) + mean_over_batch*(1/(i+1))."""
    cfg = cfg or PPConfig()
    path = pipeline.InputPath(cfg, device=device, data_mean=None, training=False)
    dev = path.device
    means = torch.zeros(9 * cfg.max_pillars * cfg.max_points_per_pillar, device=dev)
    for i, sweeps in enumerate(sample_batches):
        pts = torch.cat([torch.as_tensor(s, dtype=torch.float32) for s in sweeps]).to(dev)
        offs = [0]
        for s in sweeps:
            offs.append(offs[-1] + int(s.shape[0]))
        x, _, _ = path.pillarize(pts, offs)                       # [B,9,P,N] with data_mean = None: the raw features
        m = torch.mean(x.reshape(x.shape[0], -1), dim=0)
        means = means * (i / (i + 1)) + m * (1 / (i + 1))
    return means.cpu()
