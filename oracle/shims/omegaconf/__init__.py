"""Stub of the two names rewards/reward.py imports (test infrastructure; see oracle/shims/README.md)."""
import types


class DictConfig(dict):
    pass


class OmegaConf:
    @staticmethod
    def create(d=None):
        return types.SimpleNamespace(**dict(d or {}))
