"""TEST INFRASTRUCTURE — CPU restatement (numpy, float64) of the post-sampling pipeline rows of SURVEY.md §8(f):
reward scaling / reduction, composition-level property sums, the validity pre-filter.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module; the product never does.

Pinning status
* `reward_scoring` follows rewards/reward.py:8-12, 51-115 and is PINNED: equal (bit for bit) to the unmodified
  `rewards.reward.Reward.scoring` driven with stub calculators (tests/test_oracle_vs_reference.py).
* `cell_length_ok` follows pipeline/filters/opt_filter.py:53-55 (one comparison; in-tree).
* `composition_property` restates the mass-fraction weighted table sums of pymatgen's `HHIModel.get_hhi_reserve` /
  `CostAnalyzer.get_cost_per_kg` and of `abundance_crust` (rewards/calculators/pymatgen/calc.py:24-45, 57-92).  pymatgen
  (pinned by env.yml as `pymatgen>=2024`) and its data tables are absent from the reference tree and this image:
  PARITY UNPINNED for the property VALUES.
* `structure_validity` restates mattergen@5bb2b39 `mattergen.evaluation.utils.structure_utils.structure_validity`
  as published (min pairwise periodic distance >= 0.5 A, volume >= 0.1 A^3, no cell edge above 40 A); the package
  is un-vendored: PARITY UNPINNED.
"""
import math

import numpy as np


# ----------------------------------------------------------------------------- rewards/reward.py
def linear_scaling(values, minv=0.0, maxv=6.0):
    """rewards/reward.py:8-12"""
    ss = (np.asarray(values, dtype=np.float64) - minv) / (maxv - minv)
    ss = ss.copy()
    ss[ss > 1.0] = 1.0
    ss[ss < 0.0] = 0.0
    return ss


def reward_scoring(raw_props, cfgs, reduce="mean"):
    """rewards/reward.py:51-115.  raw_props: list (one per property, same order as cfgs) of float arrays that may hold
    NaN (failed calculator); cfgs: dicts with name / target ('ascending' | 'descending' | float) / minv / maxv
    [/ weight].  Returns (rewards, prop_dict, failed_mask)."""
    prop_list = np.array([np.asarray(p, dtype=np.float64) for p in raw_props])
    failed = np.isnan(prop_list).any(axis=0)
    prop_dict = {c["name"]: np.nan_to_num(np.asarray(p, dtype=np.float64), nan=0.0).astype(float) for c, p in zip(cfgs, raw_props)}
    scaled = {}
    for c in cfgs:
        v = prop_dict[c["name"]]
        if c["target"] == "ascending":
            s = linear_scaling(v, c["minv"], c["maxv"])
        elif c["target"] == "descending":
            s = linear_scaling(-v, -c["maxv"], -c["minv"])
        elif isinstance(c["target"], float):
            s = linear_scaling(-np.abs(v - c["target"]), -c["maxv"], -c["minv"])
        else:
            raise TypeError("prop cfg.target must be a float or descending or ascending")
        scaled[c["name"]] = s
    vals = list(scaled.values())
    if reduce == "mean":
        tot = vals[0].copy()
        for s in vals[1:]:
            tot += s
        rewards = tot / len(vals)
    elif reduce == "min":
        rewards = np.array(vals).min(axis=0)
    elif reduce == "weight":
        rewards = np.array([s * c["weight"] for s, c in zip(vals, cfgs)]).sum(axis=0)
    else:
        raise ValueError(reduce)
    rewards = rewards.copy()
    rewards[failed] = 0.0
    return rewards, prop_dict, failed


# ----------------------------------------------------------------------------- composition properties
def composition_property(atomic_numbers, table, mass, weights="mass"):
    """sum over the elements of one crystal (increasing Z) of w_el * table[el]; w = mass fraction or atom fraction"""
    z = np.asarray(atomic_numbers, dtype=np.int64)
    els, cnt = np.unique(z, return_counts=True)
    mtot = 0.0
    for e, c in zip(els, cnt):
        mtot += float(c) * float(mass[e])
    v = 0.0
    for e, c in zip(els, cnt):
        w = (float(c) * float(mass[e])) / mtot if weights == "mass" else float(c) / float(len(z))
        v += w * float(table[e])
    return v


# ----------------------------------------------------------------------------- validity pre-filter
def lattice_matrix(lengths, angles):
    """models/diffcsp/utils.py:68-96 (pymatgen Lattice.from_parameters convention), float64"""
    a, b, c = (float(v) for v in lengths)
    al, be, ga = (math.radians(float(v)) for v in angles)
    val = (math.cos(al) * math.cos(be) - math.cos(ga)) / (math.sin(al) * math.sin(be))
    gs = math.acos(max(-1.0, min(1.0, val)))
    return np.array([[a * math.sin(be), 0.0, a * math.cos(be)],
                     [-b * math.sin(al) * math.cos(gs), b * math.sin(al) * math.sin(gs), b * math.cos(al)],
                     [0.0, 0.0, c]])


def cell_length_ok(lengths, max_len=25.0):
    """pipeline/filters/opt_filter.py:53-55"""
    return bool(max(float(v) for v in lengths) < max_len)


def min_periodic_distance(frac, L):
    """minimum over atom pairs (and an atom's own images) of the image distance, 27 images around the wrapped
    fractional difference"""
    frac = np.asarray(frac, dtype=np.float64)
    n = len(frac)
    shifts = np.array([[i, j, k] for i in (-1, 0, 1) for j in (-1, 0, 1) for k in (-1, 0, 1)], dtype=np.float64)
    best = np.inf
    for j in range(n):
        for k in range(j + 1):
            d = frac[j] - frac[k]
            d = d - np.rint(d)
            for s in shifts:
                if j == k and not s.any():
                    continue
                v = (d + s) @ L
                best = min(best, float(np.sqrt((v * v).sum())))
    return best


def structure_validity(frac, lengths, angles, cutoff=0.5, min_vol=0.1, hard_len=40.0):
    L = lattice_matrix(lengths, angles)
    return bool(min_periodic_distance(frac, L) >= cutoff and abs(np.linalg.det(L)) >= min_vol and
                max(float(v) for v in lengths) <= hard_len)
