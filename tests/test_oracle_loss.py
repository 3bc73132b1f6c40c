"""The loss oracle (oracle/loss.py, fp64 restatement of model/loss.py:24-63) against the golden fixture
generated from the reference's own PPLoss module, autograd gradients included
(tests/golden/make_golden_loss.py -> loss_small.npz)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import loss as ol


@pytest.mark.parametrize("tag", ["cfg", "ort", "g3"])
def test_oracle_matches_reference_module(tag):
    g = np.load(os.path.join(GOLDEN, "loss_small.npz"))
    b_ort, b_reg, b_cls, gamma = g[tag + "/params"]
    o = ol.pp_loss(g[tag + "/cls"], g[tag + "/reg"], g[tag + "/cls_t"], g[tag + "/reg_t"], b_ort, b_reg, b_cls, gamma)
    want = g[tag + "/losses"]
    got = np.array([o["cls_loss"], o["reg_loss"], o["ort_loss"], o["total"]])
    assert np.allclose(got, want, rtol=1e-13, atol=0)
    assert np.abs(o["p"] - g[tag + "/p"]).max() < 1e-15
    assert np.abs(o["grad_cls"] - g[tag + "/grad_cls"]).max() < 1e-15
    assert np.abs(o["grad_reg"] - g[tag + "/grad_reg"]).max() < 1e-15
    assert np.abs(o["reg_after"] - g[tag + "/reg_after"]).max() < 1e-15      # the in-place tanh on channel 6 only


def test_oracle_without_positives_is_nan_like_torch():
    rng = np.random.default_rng(0)
    cls = rng.normal(0, 1, (1, 54, 3, 4)); reg = rng.normal(0, 1, (1, 48, 3, 4))
    o = ol.pp_loss(cls, reg, np.zeros((1, 72, 9)), np.zeros((1, 72, 9)), 0, 1, 250, 2)
    assert np.isnan(o["reg_loss"]) and np.isnan(o["total"]) and np.isfinite(o["cls_loss"])
    assert not o["grad_reg"].any()
