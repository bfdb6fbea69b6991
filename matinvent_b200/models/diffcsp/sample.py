"""Sampler front-end of the DiffCSP back-end (mirror of models/diffcsp/sample.py:117-201).

`DiffCSPSampler.generate(model, batch_size, num_batches, **kwargs)` draws atom counts from the mp_20
prior, runs the reverse diffusion on the GPU, and post-processes ON DEVICE (argmax atom types + 1,
lattice -> lengths/angles) before a single D2H copy.  It returns `(data_list, struc_list)` like the
reference; `data_list` items are `CrystalData` (attribute-compatible with the PyG `Data` the reference
builds, sample.py:186-193) and `struc_list` holds pymatgen `Structure`s when pymatgen is importable,
else the same `CrystalData` objects (rewards that need pymatgen are out of scope).
Unlike the reference, which keeps only the LAST loader batch (sample.py:166-177), all `num_batches`
batches are returned.
"""
from dataclasses import dataclass
from typing import List, Tuple

import numpy as np
import torch

from ... import ops
from .diffusion import DiffCSPModule

# atom-count prior (index = number of atoms), models/diffcsp/sample.py:42-62
ATOM_DIST = {
    "perov_5": [0, 0, 0, 0, 0, 1],
    "mp_20": [0.0, 0.0021742334905660377, 0.021079009433962265, 0.019826061320754717, 0.15271226415094338,
              0.047132959905660375, 0.08464770047169812, 0.021079009433962265, 0.07808814858490566,
              0.03434551886792453, 0.0972877358490566, 0.013303360849056603, 0.09669811320754718,
              0.02155807783018868, 0.06522700471698113, 0.014372051886792452, 0.06703272405660378,
              0.00972877358490566, 0.053176591981132074, 0.010576356132075472, 0.08995430424528301],
}

DEFAULT_STEP_LR = {"gen": {"perov_5": 1e-6, "carbon_24": 1e-5, "mp_20": 5e-6}}   # sample.py:66-84


class CrystalData:
    """One sampled crystal; attribute names follow the PyG Data of sample.py:186-193."""

    def __init__(self, frac_coords, atom_types, lengths, angles, num_atoms):
        self.frac_coords, self.atom_types = frac_coords, atom_types
        self.lengths, self.angles = lengths, angles
        self.num_atoms = num_atoms
        self.num_nodes = num_atoms
        self.reward = None

    def to(self, *a, **k):
        return self


class CrystalBatch:
    """Collated crystals (what the reference gets from the PyG DataLoader)."""

    def __init__(self, data_list, device=None):
        self.num_graphs = len(data_list)
        self.num_atoms = torch.tensor([int(d.num_atoms) for d in data_list], dtype=torch.int64)
        self.num_nodes = int(self.num_atoms.sum())
        self.batch = torch.repeat_interleave(torch.arange(self.num_graphs), self.num_atoms)
        cat = lambda k: torch.cat([getattr(d, k) for d in data_list])
        if data_list and data_list[0].frac_coords is not None:
            self.frac_coords = cat("frac_coords")
            self.atom_types = cat("atom_types")
            self.lengths, self.angles = cat("lengths"), cat("angles")
        if data_list and getattr(data_list[0], "reward", None) is not None:
            self.reward = torch.cat([d.reward.reshape(-1) for d in data_list])
        if device is not None:
            self.to(device)

    def to(self, device):
        for k, v in list(vars(self).items()):
            if torch.is_tensor(v):
                setattr(self, k, v.to(device))
        return self


class SampleDataset:
    """Atom counts drawn from the dataset prior with numpy's global RNG (sample.py:117-138)."""

    def __init__(self, total_num, dataset="mp_20", num_atoms=None):
        self.total_num = total_num
        self.distribution = ATOM_DIST[dataset]
        if num_atoms is not None:          # a shard of a draw made elsewhere (multi-GPU sampling)
            assert len(num_atoms) == total_num
            self.num_atoms = np.asarray(num_atoms, dtype=np.int64)
        else:
            self.num_atoms = np.random.choice(len(self.distribution), total_num, p=self.distribution)

    def __len__(self):
        return self.total_num

    def batches(self, batch_size):
        for i in range(0, self.total_num, batch_size):
            na = self.num_atoms[i:i + batch_size]
            yield CrystalBatch([CrystalData(None, None, None, None, int(n)) for n in na])


_PYMATGEN = None          # (Lattice, Structure), False when pymatgen is absent; resolved once (a failing import per crystal
                          # costs 0.2 ms: 45 ms per 256-crystal batch)


def to_structure(data):
    """pymatgen Structure of a sampled crystal (sample.py:87-100) when pymatgen is installed."""
    global _PYMATGEN
    if _PYMATGEN is None:
        try:
            from pymatgen.core.lattice import Lattice
            from pymatgen.core.structure import Structure
            _PYMATGEN = (Lattice, Structure)
        except ImportError:
            _PYMATGEN = False
    if not _PYMATGEN:
        return data
    Lattice, Structure = _PYMATGEN
    lat = Lattice.from_parameters(*(data.lengths[0].tolist() + data.angles[0].tolist()))
    return Structure(lattice=lat, species=data.atom_types.numpy(), coords=data.frac_coords.numpy(),
                     coords_are_cartesian=False)


def postprocess(outputs):
    """argmax(atom types)+1 and lattice -> (lengths, angles) on the device, then one D2H copy per field
    and the per-crystal split (sample.py:174-199)."""
    a, l = outputs["atom_types"].contiguous(), outputs["lattices"].contiguous()
    N, B = a.shape[0], l.shape[0]
    types = torch.empty(N, dtype=torch.int32, device=a.device)
    ops.argmax_rows(a, N, a.shape[1], types, add=1)
    lengths, angles = torch.empty(B, 3, device=a.device), torch.empty(B, 3, device=a.device)
    ops.lattice_matrix_to_params(l, lengths, angles, B)
    x = outputs["frac_coords"].cpu()
    types, lengths, angles = types.cpu().to(torch.int64), lengths.cpu(), angles.cpu()
    na = outputs["num_atoms"].cpu()
    off = [0] + torch.cumsum(na, 0).tolist()
    data = []
    for i in range(B):
        data.append(CrystalData(x[off[i]:off[i + 1]], types[off[i]:off[i + 1]], lengths[i].view(1, -1),
                                angles[i].view(1, -1), na[i]))
    return data


def pack_crystals(data_list):
    """a list of CrystalData as five concatenated CPU tensors (one pickle of five tensors instead of five per crystal when
    sampled crystals travel between ranks)"""
    if not data_list:
        return None
    return (torch.cat([d.frac_coords for d in data_list]), torch.cat([d.atom_types for d in data_list]),
            torch.cat([d.lengths for d in data_list]), torch.cat([d.angles for d in data_list]),
            torch.tensor([int(d.num_atoms) for d in data_list], dtype=torch.int64))


def unpack_crystals(pack):
    if pack is None:
        return []
    x, z, lengths, angles, na = pack
    off = [0] + torch.cumsum(na, 0).tolist()
    return [CrystalData(x[off[i]:off[i + 1]], z[off[i]:off[i + 1]], lengths[i].view(1, -1), angles[i].view(1, -1), na[i])
            for i in range(len(na))]


@dataclass
class DiffCSPSampler:
    batch_size: int | None = None
    num_batches: int | None = None
    target_compositions_dict: list | None = None
    num_atoms_distribution: str = "mp_20"

    def generate(self, model: DiffCSPModule, batch_size=None, num_batches=None, noise=None, num_atoms=None,
                 **kwargs) -> Tuple[List, List]:
        """`num_atoms` (optional): the atom counts to sample instead of drawing them (a rank's shard of the global draw);
        further keyword arguments (`filter`, `max_num`, `mlip_opt` ... of the pipeline's sample_cfg) are ignored like
        the reference's `**kwargs`."""
        batch_size = batch_size or self.batch_size
        num_batches = num_batches or self.num_batches
        assert batch_size is not None and num_batches is not None
        model.eval()
        dataset = SampleDataset(batch_size * num_batches, self.num_atoms_distribution, num_atoms=num_atoms)
        step_lr = DEFAULT_STEP_LR["gen"][self.num_atoms_distribution]
        data_list = []
        for batch in dataset.batches(batch_size):
            outputs, _ = model.sample(batch, step_lr=step_lr, noise=noise)
            data_list += postprocess(outputs)
        return data_list, [to_structure(d) for d in data_list]
