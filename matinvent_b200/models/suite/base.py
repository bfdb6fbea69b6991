"""Plugin API of the model back-ends — mirror of models/suite/base.py:30-59 (same constructor, same four
methods); this is the drop-in boundary `pipeline/base.py` talks to."""
import torch

from ...config import Config


def get_device(device=None):
    if device is None:
        if not torch.cuda.is_available():
            raise RuntimeError("matinvent_b200 needs a CUDA device (no CPU fallback)")
        device = "cuda"
    return torch.device(device)


class ModelSuite:
    def __init__(self, model_name, sample_cfg, finetune_cfg, model_path=None, config_overrides=(), device=None,
                 **kwargs):
        self.model_name = model_name
        self.sample_cfg = Config(sample_cfg)
        self.finetune_cfg = Config(finetune_cfg)
        self.model_path = model_path
        self.config_overrides = list(config_overrides)
        self.device = get_device(device)
        self.cfg = Config(kwargs)

    def load_model(self):
        raise NotImplementedError

    def get_sampler(self):
        raise NotImplementedError

    def get_dataloader(self):
        raise NotImplementedError

    def save_model(self):
        raise NotImplementedError
