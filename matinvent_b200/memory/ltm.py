"""Long-term memory with the diversity filter — same behaviour as memory/ltm.py:8-188 (`LongTimeMem`), without
pandas and without pymatgen strings.

Every generated crystal is remembered by two 64-bit keys computed on the device by `mi_composition_key` (the hash
the replay buffer already uses): the gcd-reduced element-count vector (the equivalence class of pymatgen's
`composition.reduced_formula`, ltm.py:31) and the element SET (the same hash of each element counted once:
`ele_comb`, ltm.py:33-37), plus its reward and RL step.  The memory lives in HBM; `div_filter` (ltm.py:65-109: the
Augmented-Hill-Climb occurrence penalty) counts occurrences with one sort of the memory and two binary searches per
sample instead of a pandas `value_counts()` per sample; `calc_metrics` / `get_baseline` / `unique_comps` are
reductions over the same arrays.  Structures themselves are kept on the host only when `keep_structures` is set
(the reference keeps them for its CSV dump).
"""
import numpy as np
import torch

from .replay_buffer import ReplayBuffer, _atomic_numbers


class LongTimeMem:
    def __init__(self, device=None, keep_structures=False):
        self.device = torch.device(device if device is not None else "cuda")
        self.keep_structures = keep_structures
        self.strucs = []
        self._keyer = None
        self.comp = torch.zeros(0, dtype=torch.int64, device=self.device)       # reduced-composition key
        self.ele = torch.zeros(0, dtype=torch.int64, device=self.device)        # element-set key
        self.reward = torch.zeros(0, dtype=torch.float64, device=self.device)
        self.rl_step = torch.zeros(0, dtype=torch.int64, device=self.device)

    # ------------------------------------------------------------------ keys (device kernel)
    def keys_of(self, strucs):
        """(composition keys, element-set keys) of pymatgen Structures or sampled crystals (anything with
        `atomic_numbers` / `atom_types`)"""
        if self._keyer is None:
            self._keyer = ReplayBuffer(device=self.device)
        z = [_atomic_numbers(s) for s in strucs]
        return self._keyer.keys_of(z), self._keyer.keys_of([torch.unique(v) for v in z])

    # ------------------------------------------------------------------ reference API
    def extend(self, strucs, rewards, step):
        """ltm.py:30-63"""
        if len(strucs) == 0:
            return
        comp, ele = self.keys_of(strucs)
        self.extend_keys(comp, ele, rewards, step)
        if self.keep_structures:
            self.strucs.extend(strucs)

    def extend_keys(self, comp, ele, rewards, step):
        dev = self.device
        r = torch.as_tensor(np.asarray(rewards, dtype=np.float64)).to(dev).reshape(-1)
        self.comp = torch.cat([self.comp, comp.to(dev, torch.int64)])
        self.ele = torch.cat([self.ele, ele.to(dev, torch.int64)])
        self.reward = torch.cat([self.reward, r])
        self.rl_step = torch.cat([self.rl_step, torch.full((r.numel(),), int(step), dtype=torch.int64, device=dev)])

    @property
    def unique_comps(self):
        """ltm.py:63 (the reference holds the array of formulas; only its length is ever read)"""
        return torch.unique(self.comp)

    def div_filter(self, strucs, rewards, tol=10, buff=20, method="composition", **kwargs):
        """ltm.py:65-109: occurrences <= tol keep the reward, tol < occ < buff scale it by (buff - occ) / (buff - tol),
        occ >= buff zero it and report the index.  Returns (new_rewards, penalty_idx, tol_n, buff_n)."""
        comp, ele = self.keys_of(strucs) if len(strucs) else (self.comp[:0], self.ele[:0])
        return self.div_filter_keys(comp if method == "composition" else ele, rewards, tol, buff, method)

    def div_filter_keys(self, keys, rewards, tol=10, buff=20, method="composition"):
        assert tol < buff
        if method not in ("composition", "element_comb"):
            raise ValueError("method must be 'composition' or 'element_comb'")
        mem = self.comp if method == "composition" else self.ele
        r = torch.as_tensor(np.asarray(rewards, dtype=np.float64)).to(self.device).reshape(-1)
        keys = keys.to(self.device, torch.int64)
        srt = torch.sort(mem).values
        occ = torch.searchsorted(srt, keys, right=True) - torch.searchsorted(srt, keys, right=False)
        soft = (occ > tol) & (occ < buff)
        hard = occ >= buff
        ramp = r * (buff - occ).double() / float(buff - tol)          # the reference's order: multiply, then divide
        new = torch.where(hard, torch.zeros_like(r), torch.where(soft, ramp, r))
        penalty_idx = torch.nonzero(hard).reshape(-1).tolist()
        return new.cpu().numpy(), penalty_idx, int(soft.sum()), int(hard.sum())

    def calc_metrics(self, thred, budget=3000, num_candidate=100):
        """ltm.py:111-133: burden = crystals generated per unique composition whose best reward exceeds `thred` (None
        below `num_candidate` such compositions); diversity ratio = unique compositions / crystals while within `budget`."""
        n = len(self)
        if n == 0:
            return None, None
        order = torch.argsort(self.reward, descending=True, stable=True)
        comp = self.comp[order]
        first = torch.ones(n, dtype=torch.bool, device=self.device)
        by_key = torch.argsort(comp, stable=True)                   # best reward first inside every key group
        ck = comp[by_key]
        first[1:] = ck[1:] != ck[:-1]
        best = self.reward[order][by_key][first]
        candidates = int((best > thred).sum())
        burden = n / candidates if candidates >= num_candidate else None
        div_ratio = int(first.sum()) / n if n <= budget else None
        return burden, div_ratio

    def get_baseline(self, step, prev=3):
        """ltm.py:135-137"""
        m = self.rl_step > step - prev
        return float(self.reward[m].mean()) if bool(m.any()) else float("nan")

    def deduplicate_indices(self):
        """ltm.py:139-149 (method='composition'): indices of the best-reward crystal of every composition"""
        order = torch.argsort(self.reward, descending=True, stable=True)
        by_key = order[torch.argsort(self.comp[order], stable=True)]
        ck = self.comp[by_key]
        first = torch.ones(len(self), dtype=torch.bool, device=self.device)
        first[1:] = ck[1:] != ck[:-1]
        return by_key[first]

    def save(self, save_path):
        """ltm.py:161-166 writes struc/comp/reward/step (+ CIF); here: key, element-set key, reward, step (+ CIF when the
        structures were kept and can write one)"""
        import csv
        cols = [self.comp.cpu().tolist(), self.ele.cpu().tolist(), self.reward.cpu().tolist(), self.rl_step.cpu().tolist()]
        with open(save_path, "w", newline="") as fh:
            w = csv.writer(fh, quoting=csv.QUOTE_ALL)
            w.writerow(["comp_key", "ele_comb_key", "reward", "RL_step", "cif"])
            for i, row in enumerate(zip(*cols)):
                s = self.strucs[i] if i < len(self.strucs) else None
                w.writerow(["%016x" % (row[0] & (2 ** 64 - 1)), "%016x" % (row[1] & (2 ** 64 - 1)), row[2], row[3],
                            s.to(fmt="cif") if hasattr(s, "to") and hasattr(s, "lattice") else ""])

    def __len__(self):
        return int(self.reward.numel())
