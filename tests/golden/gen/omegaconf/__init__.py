"""Minimal stand-in for the `omegaconf` package, used ONLY by the golden-vector
generator (tests/golden/gen/make_goldens.py) so that the unmodified reference
(`/root/reference/pixloc`, `/root/reference/pixtrack`) can be imported in the
authoring container, where omegaconf is not installed and there is no network.

It is never imported by the product package, by the oracle, or by any test.
It implements just what the reference touches on the optimizer / UNet path:
OmegaConf.create / merge / set_struct / set_readonly / to_yaml, attribute +
item access on nested dict configs, `read_write` and `open_dict` contexts.
"""
import contextlib
import copy


class DictConfig(dict):
    def __init__(self, src=None):
        super().__init__()
        for k, v in (src or {}).items():
            dict.__setitem__(self, k, _wrap(v))

    def __getattr__(self, key):
        try:
            return self[key]
        except KeyError as exc:
            raise AttributeError(key) from exc

    def __setattr__(self, key, value):
        self[key] = _wrap(value)

    def get(self, key, default=None):
        return self[key] if key in self else default

    def pop(self, key, *default):
        return dict.pop(self, key, *default)

    def __deepcopy__(self, memo):
        return DictConfig({k: copy.deepcopy(v, memo) for k, v in self.items()})


class ListConfig(list):
    pass


def _wrap(v):
    if isinstance(v, DictConfig):
        return v
    if isinstance(v, dict):
        return DictConfig(v)
    if isinstance(v, (list, tuple)) and not isinstance(v, ListConfig):
        return ListConfig(_wrap(x) for x in v)
    return v


def _merge_into(dst, src):
    for k, v in src.items():
        if isinstance(v, dict) and isinstance(dst.get(k), dict):
            _merge_into(dst[k], v)
        else:
            dict.__setitem__(dst, k, copy.deepcopy(_wrap(v)))


class OmegaConf:
    @staticmethod
    def create(src=None):
        return DictConfig(copy.deepcopy(src) if src else {})

    @staticmethod
    def merge(*confs):
        out = DictConfig()
        for c in confs:
            _merge_into(out, c if isinstance(c, dict) else {})
        return out

    @staticmethod
    def set_struct(conf, flag):
        return None

    @staticmethod
    def set_readonly(conf, flag):
        return None

    @staticmethod
    def to_yaml(conf):
        return repr(dict(conf))


@contextlib.contextmanager
def read_write(conf):
    yield conf


@contextlib.contextmanager
def open_dict(conf):
    yield conf
