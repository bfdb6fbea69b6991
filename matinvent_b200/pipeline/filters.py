"""Validity pre-filter of sampled crystals on the device — mirror of pipeline/filters/opt_filter.py:50-63
(`invalid_filter`): mask = structure_validity & is_smact_valid & (max(a, b, c) < 25 A).

`mi_validity_prefilter` evaluates, one warp per crystal, the cell-length rule (:53-55, in-tree, pinned) and the
geometric predicate of mattergen's `structure_validity` (:51; un-vendored package — restated from its published
definition: minimum periodic interatomic distance >= 0.5 A, cell volume >= 0.1 A^3, no cell edge above 40 A; parity
unpinned) for the whole batch in one launch instead of a Python loop over pymatgen lattices inside an `mp.Pool`.
`is_smact_valid` (:52) needs SMACT's oxidation-state tables, absent from this image: it is a hook (callable over the
structure list) and defaults to "valid"."""
import numpy as np
import torch

from .. import ops
from ..models.diffcsp.sample import CrystalBatch

MAX_CELL_LENGTH = 25.0


def validity_masks(sample_data, device=None, max_len=MAX_CELL_LENGTH, min_dist=0.5, min_vol=0.1, hard_len=40.0):
    """(cell_ok [B] bool, structure_ok [B] bool, min_distance [B]) numpy arrays for a list of sampled crystals"""
    B = len(sample_data)
    if B == 0:
        z = np.zeros(0, dtype=bool)
        return z, z.copy(), np.zeros(0, dtype=np.float32)
    dev = torch.device(device if device is not None else "cuda")
    cb = CrystalBatch(sample_data)
    frac = cb.frac_coords.to(dev, torch.float32).contiguous()
    lengths = cb.lengths.to(dev, torch.float32).reshape(B, 3).contiguous()
    angles = cb.angles.to(dev, torch.float32).reshape(B, 3).contiguous()
    off = torch.zeros(B + 1, dtype=torch.int32)
    off[1:] = torch.cumsum(cb.num_atoms, 0).to(torch.int32)
    off = off.to(dev)
    L = torch.empty(B, 3, 3, device=dev)
    ops.lattice_params_to_matrix(lengths, angles, L, B)
    mask = torch.empty(B, dtype=torch.int32, device=dev)
    dmin = torch.empty(B, device=dev)
    ops.validity_prefilter(frac, L, lengths, off, B, mask, dmin, max_len, min_dist, min_vol, hard_len)
    m = mask.cpu().numpy()
    return (m & 1).astype(bool), (m & 2).astype(bool), dmin.cpu().numpy()


def cell_length_mask(sample_data, max_len=MAX_CELL_LENGTH, device=None):
    """mask[i] = max(lengths_i) < max_len   (opt_filter.py:53-55)"""
    return validity_masks(sample_data, device=device, max_len=max_len)[0]


def invalid_filter(sample_data, sample_struc, return_mask=False, structure_validity=True, smact_validity=None, device=None,
                   **thresholds):
    """opt_filter.py:50-63.  structure_validity: True = the device predicate, a callable = the caller's own (e.g.
    mattergen's), None/False = skipped; smact_validity: a callable over structures or None (skipped); thresholds
    (max_len, min_dist, min_vol, hard_len) default to the reference's."""
    cell_ok, struc_ok, _ = validity_masks(sample_data, device=device, **thresholds)
    mask = cell_ok.copy()
    if structure_validity is True:
        mask &= struc_ok
    elif callable(structure_validity):
        mask &= np.asarray([bool(structure_validity(s)) for s in sample_struc], dtype=bool)
    if smact_validity is not None:
        mask &= np.asarray([bool(smact_validity(s)) for s in sample_struc], dtype=bool)
    if return_mask:
        return mask
    return ([x for x, m in zip(sample_data, mask) if m], [x for x, m in zip(sample_struc, mask) if m])
