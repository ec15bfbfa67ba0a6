"""Minimal attribute-style dict with the part of easydict 1.10 the reference uses
(utils/config.py:50-66: EasyDict(json_dict), attribute get/set)."""


class EasyDict(dict):
    def __init__(self, d=None, **kwargs):
        super().__init__()
        merged = dict(d or {})
        merged.update(kwargs)
        for k, v in merged.items():
            setattr(self, k, v)

    def __setattr__(self, name, value):
        if isinstance(value, dict) and not isinstance(value, EasyDict):
            value = EasyDict(value)
        super().__setitem__(name, value)

    __setitem__ = __setattr__

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError as e:
            raise AttributeError(name) from e
