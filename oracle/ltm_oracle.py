"""TEST INFRASTRUCTURE — CPU restatement of the bookkeeping of the reference's long-term memory (memory/ltm.py:8-188)
on plain Python containers; any hashable stands in for pymatgen's `composition.reduced_formula` / element tuple (the
pymatgen calls that produce those keys are not restated).  Only tests may import this.

What the reference does, per method:
  extend        (ltm.py:30-63)    append (comp, ele_comb, reward, RL_step) rows; unique_comps = distinct comps
  div_filter    (ltm.py:65-109)   occ = rows of the memory carrying the sample's key; occ <= tol: reward kept;
                                  tol < occ < buff: reward * (buff - occ) / (buff - tol); otherwise 0 and index reported
  calc_metrics  (ltm.py:111-133)  burden = rows / #(distinct comps whose best reward > thred), None below num_candidate;
                                  div_ratio = distinct comps / rows while rows <= budget, else None
  get_baseline  (ltm.py:135-137)  mean reward of the rows with RL_step > step - prev
"""
from collections import Counter

import numpy as np


class LongTimeMemOracle:
    def __init__(self):
        self.rows = []                      # (comp, ele_comb, reward, step)
        self.unique_comps = []

    @property
    def memory(self):
        return self.rows

    def extend(self, comps, ele_comb, rewards, step):
        self.rows += [(c, e, float(r), step) for c, e, r in zip(comps, ele_comb, rewards)]
        self.unique_comps = sorted({c for c, _, _, _ in self.rows}, key=repr)

    def div_filter(self, values, rewards, tol=10, buff=20, method="composition"):
        if not tol < buff:
            raise AssertionError
        col = 0 if method == "composition" else 1
        seen = Counter(row[col] for row in self.rows)
        occ = np.array([seen.get(v, 0) for v in values], dtype=np.int64)
        rewards = np.asarray(rewards, dtype=float)
        out = np.where(occ <= tol, rewards, np.where(occ < buff, rewards * (buff - occ) / (buff - tol), 0.0))    # multiply, then divide
        hard = occ >= buff
        return out, np.flatnonzero(hard).tolist(), int(((occ > tol) & ~hard).sum()), int(hard.sum())

    def calc_metrics(self, thred, budget=3000, num_candidate=100):
        best = {}
        for c, _, r, _ in self.rows:
            best[c] = max(r, best.get(c, -np.inf))
        good = sum(1 for r in best.values() if r > thred)
        n = len(self.rows)
        return (n / good if good >= num_candidate else None), (len(best) / n if n <= budget else None)

    def get_baseline(self, step, prev=3):
        recent = [r for _, _, r, s in self.rows if s > step - prev]
        return float(np.mean(recent)) if recent else float("nan")

    def best_row_per_comp(self):
        """ltm.py:139-149 (method='composition'): row index of the highest reward of every comp (first one on ties, as a
        stable descending sort followed by drop_duplicates keeps)"""
        best = {}
        for i, (c, _, r, _) in enumerate(self.rows):
            if c not in best or r > self.rows[best[c]][2]:
                best[c] = i
        return sorted(best.values())
