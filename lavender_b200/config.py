"""Config surface: an EasyDict-compatible `Args` (utils/lib.py:5 `edict`) and the JSON-over-defaults merge of
utils/args.py:16-34 for the hot-path keys.  The reference's own utils/args.py runs unchanged on top of the drop-in
`utils/lib.py`; this module is what bench.py / tests use without the reference tree."""
import json


class Args(dict):
    """dict with attribute access; missing attributes raise AttributeError (model.py:11-13 relies on
    getattr(args, 'swinbert', False))."""

    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        for k, v in list(self.items()):
            if isinstance(v, dict) and not isinstance(v, Args):
                self[k] = Args(v)

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v

    def __delattr__(self, k):
        try:
            del self[k]
        except KeyError:
            raise AttributeError(k)


def load_args(path, **overrides):
    """_args/*.json -> Args; keyword overrides win (the CLI-over-JSON precedence of utils/args.py:16-34)."""
    with open(path) as f:
        a = Args(json.load(f))
    a.update(overrides)
    return a
