"""Minimal attribute-style config tree (OmegaConf is not a dependency of the hot path).

Covers what pipeline/base.py:53-59 and models/suite/*.py need: attribute + item access, recursive merge,
yaml load/save and `${a.b}` interpolation of hparams.yaml files."""
import re

import yaml


class Config(dict):
    def __init__(self, d=None, **kw):
        super().__init__()
        for k, v in dict(d or {}, **kw).items():
            self[k] = v

    def __setitem__(self, k, v):
        if isinstance(v, dict) and not isinstance(v, Config):
            v = Config(v)
        elif isinstance(v, (list, tuple)):
            v = [Config(x) if isinstance(x, dict) and not isinstance(x, Config) else x for x in v]
        super().__setitem__(k, v)

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v

    def to_container(self):
        def conv(v):
            if isinstance(v, Config):
                return {k: conv(x) for k, x in v.items()}
            if isinstance(v, list):
                return [conv(x) for x in v]
            return v
        return conv(self)

    @staticmethod
    def merge(*cfgs):
        out = Config()
        for c in cfgs:
            for k, v in (c or {}).items():
                if isinstance(v, dict) and isinstance(out.get(k), dict):
                    out[k] = Config.merge(out[k], v)
                else:
                    out[k] = v
        return out

    @staticmethod
    def load(path):
        with open(path) as fh:
            return Config(yaml.safe_load(fh)).resolve()

    def save(self, path):
        with open(path, "w") as fh:
            yaml.safe_dump(self.to_container(), fh, sort_keys=False)

    def resolve(self, root=None):
        """Replace `${a.b.c}` strings by the referenced value (whole-string references keep their type)."""
        root = root if root is not None else self
        pat = re.compile(r"\$\{([A-Za-z0-9_.]+)\}")

        def lookup(path):
            cur = root
            for part in path.split("."):
                cur = cur[part]
            return cur

        def res(v, depth=0):
            if isinstance(v, Config):
                for k in list(v.keys()):
                    v[k] = res(v[k], depth)
                return v
            if isinstance(v, list):
                return [res(x, depth) for x in v]
            if isinstance(v, str) and depth < 8:
                m = pat.fullmatch(v)
                try:
                    if m:
                        return res(lookup(m.group(1)), depth + 1)
                    return pat.sub(lambda mm: str(lookup(mm.group(1))), v)
                except (KeyError, TypeError):
                    return v
            return v
        return res(self)
