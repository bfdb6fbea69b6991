"""Validity pre-filter of sampled crystals — mirror of pipeline/filters/opt_filter.py:50-63 (`invalid_filter`).

Of its three predicates only the cell-length rule (`max(a, b, c) < 25 A`, :53-55) is arithmetic on the sampler's
output; it is evaluated on the post-processed lengths (device or host tensors, one comparison for the whole batch
instead of a Python loop over pymatgen lattices).  The composition / structure validity predicates
(`structure_validity`, `is_smact_valid`, :51-52) come from pymatgen / smact, absent from this image: they are hooks
(callables over the structure list) and default to "valid"."""
import numpy as np
import torch

MAX_CELL_LENGTH = 25.0


def cell_length_mask(sample_data, max_len=MAX_CELL_LENGTH):
    """mask[i] = max(lengths_i) < max_len   (opt_filter.py:53-55)"""
    if len(sample_data) == 0:
        return np.zeros(0, dtype=bool)
    lengths = torch.stack([torch.as_tensor(d.lengths).reshape(3) for d in sample_data])
    return (lengths.amax(dim=1) < max_len).cpu().numpy()


def invalid_filter(sample_data, sample_struc, return_mask=False, structure_validity=None, smact_validity=None):
    mask = cell_length_mask(sample_data)
    for pred in (structure_validity, smact_validity):
        if pred is not None:
            mask &= np.asarray([bool(pred(s)) for s in sample_struc], dtype=bool)
    if return_mask:
        return mask
    return ([x for x, m in zip(sample_data, mask) if m], [x for x, m in zip(sample_struc, mask) if m])
