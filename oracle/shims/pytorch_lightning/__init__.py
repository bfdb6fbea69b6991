"""Shim: LightningModule = nn.Module + dict-like hparams + .device."""
import torch


class _HP(dict):
    __getattr__ = dict.__getitem__

    def __setattr__(self, k, v):
        self[k] = v


class LightningModule(torch.nn.Module):
    def __init__(self, *a, **k):
        super().__init__()
        self._hp = _HP()

    @property
    def hparams(self):
        return self._hp

    def save_hyperparameters(self, *a, **k):
        pass

    @property
    def device(self):
        return next(self.parameters()).device
