"""Stand-in for the ``easydict`` package (absent from this image): attribute access on a dict,
recursively for nested dicts, which is all the reference's config.py uses.  TEST INFRASTRUCTURE ONLY."""


class EasyDict(dict):
    def __init__(self, d=None, **kwargs):
        super().__init__()
        d = dict(d or {})
        d.update(kwargs)
        for k, v in d.items():
            setattr(self, k, v)

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        if isinstance(v, dict) and not isinstance(v, EasyDict):
            v = EasyDict(v)
        self[k] = v
