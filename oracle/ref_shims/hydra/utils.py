import functools
import importlib


def _resolve(path):
    mod, _, name = path.rpartition(".")
    return getattr(importlib.import_module(mod), name)


def instantiate(cfg, *args, **kwargs):
    cfg = dict(cfg)
    target = _resolve(cfg.pop("_target_"))
    partial = cfg.pop("_partial_", False)
    kw = {}
    for k, v in cfg.items():
        if isinstance(v, dict) and "_target_" in v:
            v = instantiate(v)
        kw[k] = v
    kw.update(kwargs)
    if partial:
        return functools.partial(target, *args, **kw)
    return target(*args, **kw)
