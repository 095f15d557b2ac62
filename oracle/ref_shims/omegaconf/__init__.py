"""Minimal stand-in for omegaconf (test-only; see README.md)."""
import copy
import dataclasses


class DictConfig(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    def __deepcopy__(self, memo):
        return DictConfig({k: copy.deepcopy(v, memo) for k, v in self.items()})


class ListConfig(list):
    pass


def _wrap(x):
    if isinstance(x, dict):
        return DictConfig({k: _wrap(v) for k, v in x.items()})
    if isinstance(x, (list, tuple)):
        return ListConfig([_wrap(v) for v in x])
    return x


class OmegaConf:
    @staticmethod
    def create(x=None):
        return _wrap(x or {})

    @staticmethod
    def structured(obj):
        if dataclasses.is_dataclass(obj):
            inst = obj() if isinstance(obj, type) else obj
            return _wrap(dataclasses.asdict(inst))
        return _wrap(obj)

    @staticmethod
    def set_struct(cfg, flag):
        return None

    @staticmethod
    def to_container(cfg, resolve=True):
        return dict(cfg)

    @staticmethod
    def to_yaml(cfg):
        return repr(cfg)
