"""MatterGen sampler front-end — mirror of models/mattergen/sample.py:27-64, 127-303.

The reference composes mattergen's hydra sampling config, instantiates its `PredictorCorrector` and keeps, for every
batch, the noise-free `mean` of the last predictor step (sample.py:49-50); in-tree are the batching, the conversion
lattice -> (lengths, angles) and the split into per-crystal records.  Here the predictor-corrector is an injected
factory (`sampler_factory(model) -> object with .sample(conditioning_data, mask) -> (sample, mean)`, i.e. what
`instantiate(sampling_config.sampler_partial)(pl_module=model)` returns where mattergen is installed); atom counts
come from the same prior the DiffCSP front-end uses (mattergen's ALEX_MP_20 histogram is not in the reference tree),
and the post-processing runs on the device (`mi_lattice_matrix_to_params`)."""
from dataclasses import dataclass
from typing import Callable

import numpy as np
import torch

from ... import ops
from ..diffcsp.sample import ATOM_DIST, CrystalData, to_structure


def draw_samples_from_sampler(sampler, condition_loader):
    """sample.py:27-64: run the sampler over the condition loader, keep the MEANS, convert cells to (lengths, angles) on
    the device, return (per-crystal records, structures)"""
    records = []
    for conditioning_data, mask in condition_loader:
        sample, mean = sampler.sample(conditioning_data, mask)
        pos, cell = mean["pos"].reshape(-1, 3), mean["cell"].reshape(-1, 3, 3).to(torch.float32).contiguous()
        z, na = mean["atomic_numbers"].reshape(-1), mean["num_atoms"].reshape(-1)
        B = cell.shape[0]
        lengths, angles = torch.empty(B, 3, device=cell.device), torch.empty(B, 3, device=cell.device)
        ops.lattice_matrix_to_params(cell, lengths, angles, B)
        pos, z, lengths, angles, na = pos.cpu(), z.cpu().to(torch.int64), lengths.cpu(), angles.cpu(), na.cpu()
        off = [0] + torch.cumsum(na, 0).tolist()
        for i in range(B):
            d = CrystalData(pos[off[i]:off[i + 1]], z[off[i]:off[i + 1]], lengths[i].view(1, 3), angles[i].view(1, 3), na[i])
            d.cell = cell[i].cpu()
            records.append(d)
    return records, [to_structure(d) for d in records]


@dataclass
class MatterGenSampler:
    batch_size: int | None = None
    num_batches: int | None = None
    num_atoms_distribution: str = "mp_20"
    sampler_factory: Callable | None = None

    def condition_loader(self, batch_size, num_batches, device):
        """number-of-atoms conditioning batches (mattergen's get_number_of_atoms_condition_loader): num_atoms drawn from the
        prior with numpy's global RNG; every batch comes with a `mask` dict (nothing is held fixed)"""
        dist = ATOM_DIST[self.num_atoms_distribution]
        for _ in range(num_batches):
            na = np.random.choice(len(dist), batch_size, p=dist)
            yield dict(num_atoms=torch.as_tensor(na, device=device)), {}

    def generate(self, model, batch_size=None, num_batches=None, **kwargs):
        batch_size = batch_size or self.batch_size
        num_batches = num_batches or self.num_batches
        assert batch_size is not None and num_batches is not None
        if self.sampler_factory is None:
            raise RuntimeError("MatterGenSampler needs a predictor-corrector: mattergen (microsoft/mattergen@5bb2b39) is not "
                               "installed here; pass sampler_factory=lambda model: <object with .sample(cond, mask)>")
        sampler = self.sampler_factory(model)
        device = next(model.parameters()).device if any(True for _ in model.parameters()) else torch.device("cuda")
        return draw_samples_from_sampler(sampler, self.condition_loader(batch_size, num_batches, device))
