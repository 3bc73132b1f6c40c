"""``LyftDataset`` is only imported by the reference on the hot path's modules (utils/box_utils.py:16,
data/dataset.py), never called by the functions the oracle runs.  TEST INFRASTRUCTURE ONLY."""


class LyftDataset:
    def __init__(self, *args, **kwargs):
        raise RuntimeError("sdk_shim: the Lyft dataset itself is not available offline")
