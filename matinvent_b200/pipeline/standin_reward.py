"""Stand-in composition reward for plumbing tests and benchmarks.

The reference's `hhi` reward is pymatgen's HHIModel.get_hhi_reserve (rewards/calculators/pymatgen/calc.py:57-73)
on its bundled hhi_data.csv — neither pymatgen nor that table is available here, so reward VALUES are
"parity unpinned" (SURVEY.md §8c).  This class keeps the reference's interface and scaling
(rewards/reward.py:68-115, descending linear scaling between minv and maxv, failed -> 0) over a documented
synthetic per-element table: hhi_reserve(Z) = 500 + 45 * ((37 * Z) mod 89), mass(Z) = 2 Z."""
import numpy as np


class StandInHHIReward:
    threshold = 0.8

    def __init__(self, minv=750.0, maxv=3250.0):
        self.minv, self.maxv = minv, maxv

    @staticmethod
    def _z(s):
        if hasattr(s, "atomic_numbers"):
            return np.asarray(s.atomic_numbers, dtype=np.int64)
        return np.asarray(s.atom_types, dtype=np.int64).reshape(-1)

    def scoring(self, strucs_and_path, label="tmp"):
        strucs, _ = strucs_and_path
        vals = []
        for s in strucs:
            z = self._z(s)
            mass = 2.0 * z
            hhi = 500.0 + 45.0 * ((37 * z) % 89)
            vals.append(float((mass * hhi).sum() / mass.sum()))
        v = np.asarray(vals)
        failed = ~np.isfinite(v)
        r = np.clip((self.maxv - v) / (self.maxv - self.minv), 0.0, 1.0)
        r[failed] = 0.0
        return r, {"hhi": v}, failed
