"""DiffCSP back-end behind the plugin API — mirror of models/suite/diffcsp.py:24-145.

`load_model()` reads the reference's checkpoint layout (a directory with `hparams.yaml` and `*.ckpt` whose
`state_dict` uses the reference's parameter names: `decoder.csp_layer_0.edge_mlp.0.weight`, the scheduler
buffers ...).  There is no network in this environment, so the HF-hub default of the reference
(models/suite/diffcsp.py:47-55) is replaced by an explicit error unless `random_init=True` is passed, which
builds the upstream-default architecture with seeded random weights (benchmarks, tests)."""
import os
from pathlib import Path

import torch

from ...config import Config
from ..diffcsp.diffusion import DiffCSPModule
from ..diffcsp.finetune import CrystalLoader, DiffCSPDataset
from ..diffcsp.sample import DiffCSPSampler
from .base import ModelSuite

DEFAULT_MODEL_CFG = dict(
    decoder=dict(hidden_dim=512, num_layers=6, max_atoms=100, act_fn="silu", dis_emb="sin", num_freqs=128,
                 edge_style="fc", cutoff=7.0, max_neighbors=20, ln=True, ip=True),
    beta_scheduler=dict(timesteps=1000, scheduler_mode="cosine"),
    sigma_scheduler=dict(timesteps=1000, sigma_begin=0.005, sigma_end=0.5),
    cost_lattice=1.0, cost_coord=1.0, cost_type=20.0, time_dim=256, latent_dim=0, timesteps=1000)


def _find_ckpt(model_path):
    ckpts = list(Path(model_path).glob("*.ckpt"))
    if not ckpts:
        return None
    for ck in ckpts:
        if "last" in ck.name:
            return str(ck)
    def epoch(ck):
        try:
            return int(ck.name.split("-")[0].split("=")[1])
        except (IndexError, ValueError):
            return -1
    return str(sorted(ckpts, key=epoch)[-1])


class DiffCSPSuite(ModelSuite):
    def load_model(self):
        if self.model_path is None:
            if not self.cfg.get("random_init", False):
                raise RuntimeError("no model_path given: the reference would download jwchen25/MatInvent:diffcsp_mp20 "
                                   "from the HF hub (models/suite/diffcsp.py:47-55); pass model_path=<dir with hparams.yaml "
                                   "+ last.ckpt> or random_init=True")
            cfg = Config(model=Config.merge(DEFAULT_MODEL_CFG, self.cfg.get("model", {})))
            model = self._build(cfg.model)
            model.decoder.reset_parameters(seed=int(self.cfg.get("seed", 0)))
            scale = float(self.cfg.get("head_scale", 1.0))
            if scale != 1.0:
                for k in ("coord_w", "lattice_w", "type_w", "type_b"):
                    model.decoder.w(k).mul_(scale)
                model.decoder.weights_changed()
        else:
            model_path = os.path.abspath(self.model_path)
            cfg = Config.load(os.path.join(model_path, "hparams.yaml"))
            model = self._build(cfg.model)
            ckpt = _find_ckpt(model_path)
            if ckpt is not None:
                blob = torch.load(ckpt, map_location="cpu", weights_only=False)
                model.load_state_dict(blob["state_dict"], strict=False)
        model.config = cfg
        return model

    def _build(self, mcfg):
        keys = ("cost_lattice", "cost_coord", "cost_type", "time_dim", "latent_dim")
        kw = {k: mcfg[k] for k in keys if k in mcfg}
        return DiffCSPModule(decoder=mcfg.decoder, beta_scheduler=mcfg.beta_scheduler,
                             sigma_scheduler=mcfg.sigma_scheduler, device=self.device, **kw)

    def get_sampler(self):
        return DiffCSPSampler(batch_size=self.sample_cfg.batch_size, num_batches=self.sample_cfg.num_batches)

    def get_dataloader(self, samples, rewards, batch_size=None, shuffle=True):
        if batch_size is None:
            batch_size = self.finetune_cfg.batch_size
        return CrystalLoader(DiffCSPDataset(samples, rewards), batch_size=batch_size, shuffle=shuffle)

    def save_model(self, model, save_dir):
        os.makedirs(save_dir, exist_ok=True)
        cfg = model.config
        torch.save({"state_dict": model.state_dict(), "config": cfg.to_container()}, os.path.join(save_dir, "last.ckpt"))
        cfg.save(os.path.join(save_dir, "hparams.yaml"))
