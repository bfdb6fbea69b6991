import torch


class DiffusionLightningModule(torch.nn.Module):
    """constructor signature of mattergen's LightningModule wrapper: holds the diffusion module"""

    def __init__(self, diffusion_module, optimizer_partial=None, scheduler_partials=None):
        super().__init__()
        self.diffusion_module = diffusion_module
        self._optimizer_partial = optimizer_partial
        self._scheduler_partials = scheduler_partials or []
