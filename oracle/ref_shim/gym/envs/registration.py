import importlib

_REGISTRY = {}


def register(id, entry_point=None, **kwargs):
    _REGISTRY[id] = (entry_point, kwargs)


def make(id, **kwargs):
    entry_point, base = _REGISTRY[id]
    if isinstance(entry_point, str):
        mod, _, attr = entry_point.partition(":")
        entry_point = getattr(importlib.import_module(mod), attr)
    kw = dict(base.get("kwargs", {}))
    kw.update(kwargs)
    return entry_point(**kw)
