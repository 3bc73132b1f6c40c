"""Host-side logic that needs no GPU: the positives-list container, the cached ctypes offset arrays."""
import ctypes

import numpy as np
import torch

from pp_b200 import _lib
from pp_b200.box_utils import Positives


def test_positives_dense_round_trip_on_cpu_tensors():
    B, A = 2, 50
    rng = np.random.default_rng(1)
    cls = torch.zeros((B * A, 9)); reg = torch.zeros((B * A, 9))
    idx = torch.tensor(np.sort(rng.choice(B * A, 7, replace=False)))
    cls[idx, torch.tensor(rng.integers(0, 9, 7))] = 1
    reg[idx] = torch.tensor(rng.normal(0, 1, (7, 9)), dtype=torch.float32)
    cap = 16                                                    # capacity larger than the list: the tail is ignored
    anchor = torch.zeros(cap, dtype=torch.int32); anchor[:7] = idx.int()
    pc = torch.full((cap, 9), 7.0); pc[:7] = cls[idx]
    pr = torch.full((cap, 9), 7.0); pr[:7] = reg[idx]
    offs = torch.tensor([0, int((idx < A).sum()), 7], dtype=torch.int32)
    pos = Positives(anchor, pc, pr, offs, B, A)
    dc, dr = pos.dense()
    assert dc.shape == (B, A, 9) and torch.equal(dc.view(-1, 9), cls) and torch.equal(dr.view(-1, 9), reg)


def test_offset_arrays_are_cached_and_correct():
    a = _lib.i64_array([0, 10, 25])
    b = _lib.i64_array((0, 10, 25))
    assert a is b and list(a) == [0, 10, 25] and isinstance(a, ctypes.Array)
    c = _lib.i64_array(np.array([0, 3], dtype=np.int32))
    assert list(c) == [0, 3] and c is not a
    for k in range(300):                                        # the cache is bounded
        _lib.i64_array([0, k])
    assert len(_lib._i64_cache) <= 257
