"""oracle/pfn_backward.py against the fixture generated from the reference's own modules with torch autograd
(tests/golden/make_golden_pfn_backward.py)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import pfn_backward as ob

KEYS = ("grad_weight", "grad_bias", "grad_bn_weight", "grad_bn_bias", "grad_x")


def _close(a, b, tol=1e-11):
    assert np.abs(a - b).max() <= tol * max(1.0, np.abs(b).max())


@pytest.mark.parametrize("tag", ["pos", "mixed"])
@pytest.mark.parametrize("mode", ["train", "eval"])
def test_pfn_backward_matches_reference_autograd(tag, mode):
    g = np.load(os.path.join(GOLDEN, "pfn_backward_small.npz"))
    sd = lambda k: g["%s/sd/%s" % (tag, k)]
    o = ob.pfn_backward(g["x"], sd("conv1.weight").reshape(64, 9), sd("conv1.bias"), sd("bn1.weight"), g[tag + "/g_feat"],
                        mode == "train", 1e-5, sd("bn1.running_mean"), sd("bn1.running_var"))
    for k in KEYS:
        _close(o[k], g["%s/%s/%s" % (tag, mode, k)])


def test_scatter_then_pfn_backward_end_to_end():
    g = np.load(os.path.join(GOLDEN, "pfn_backward_small.npz"))
    gc = np.zeros((2, 64, 600, 600))
    i = g["e2e/g_canvas_index"]
    gc[i[:, 0], i[:, 1], i[:, 2], i[:, 3]] = g["e2e/g_canvas_value"]
    g_feat = ob.scatter_backward(gc, g["inds"])
    assert not g_feat[:, :, 40:].any()                     # flag-0 rows get nothing, although (0,0) carries gradient
    sd = lambda k: g["pos/sd/" + k]
    o = ob.pfn_backward(g["x"], sd("conv1.weight").reshape(64, 9), sd("conv1.bias"), sd("bn1.weight"), g_feat, True)
    for k in KEYS[:4]:
        _close(o[k], g["e2e/" + k])
