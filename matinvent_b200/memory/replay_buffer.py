"""Device-resident replay buffer — same behaviour as memory/replay_buffer.py:11-104, without pandas.

Rows live in padded struct-of-arrays tensors in HBM (frac [cap, M, 3], Z [cap, M], lengths/angles [cap, 3],
n [cap], reward [cap], key [cap]); the dedupe key is a 64-bit hash of the gcd-reduced element-count vector,
i.e. the equivalence class of pymatgen's `composition.reduced_formula` (replay_buffer.py:38) computed by
`mi_composition_key`; `extend` = concat -> sort by reward desc (stable) -> first occurrence per key ->
head(buffer_size) -> reward > cutoff, done by the single-CTA `mi_replay_select` kernel on float64 rewards; `sample` draws
min(len, sample_size) rows uniformly without replacement with numpy's global RNG (what DataFrame.sample
uses); `memory_purge` drops rows whose key matches a penalised structure."""
import numpy as np
import torch

from .. import ops
from ..models.diffcsp.sample import CrystalBatch, CrystalData


def _atomic_numbers(s):
    if hasattr(s, "atomic_numbers"):           # pymatgen Structure
        return torch.as_tensor(list(s.atomic_numbers), dtype=torch.int32)
    return torch.as_tensor(s.atom_types).to(torch.int32).reshape(-1)


class ReplayBuffer:
    MAX_SELECT = 16384          # rows mi_replay_select sorts in one CTA

    def __init__(self, buffer_size=100, sample_size=8, reward_cutoff=0.0, device=None, max_atoms=None):
        self.buffer_size, self.sample_size, self.reward_cutoff = int(buffer_size), int(sample_size), float(reward_cutoff)
        self.device = torch.device(device if device is not None else "cuda")
        self.M = max_atoms or 0
        self._rows = None          # dict of device tensors
        self.count = 0

    # ------------------------------------------------------------------ packing
    def _pack(self, data, rewards):
        """padded SoA rows of a list of crystals: concatenated on the host, ONE H2D copy per field, scattered into the
        padded layout on the device (no per-crystal device copies: 10 k crystals per iteration in BASELINE configs[4])"""
        n = len(data)
        dev = self.device
        cb = CrystalBatch(data)
        na = cb.num_atoms.to(torch.int64)
        M = max(self.M, int(na.max()))
        row = torch.repeat_interleave(torch.arange(n), na).to(dev)
        col = (torch.arange(int(na.sum())) - torch.repeat_interleave(torch.cumsum(na, 0) - na, na)).to(dev)
        frac = torch.zeros(n, M, 3, device=dev)
        Z = torch.zeros(n, M, dtype=torch.int32, device=dev)
        frac[row, col] = cb.frac_coords.to(dev, torch.float32)
        Z[row, col] = cb.atom_types.to(dev, torch.int32)
        lengths = cb.lengths.to(dev, torch.float32).reshape(n, 3).contiguous()
        angles = cb.angles.to(dev, torch.float32).reshape(n, 3).contiguous()
        rew = torch.as_tensor(np.asarray(rewards, dtype=np.float64)).to(dev).reshape(n)     # float64, like the pandas column
        return dict(frac=frac, Z=Z, lengths=lengths, angles=angles, n=na.to(dev, torch.int32), reward=rew)

    def keys_of(self, atomic_number_lists):
        """64-bit composition keys of a list of per-crystal atomic-number vectors."""
        counts = [int(z.numel()) for z in atomic_number_lists]
        if not counts:
            return torch.zeros(0, dtype=torch.int64, device=self.device)
        off = torch.tensor([0] + np.cumsum(counts).tolist(), dtype=torch.int32, device=self.device)
        Z = torch.cat([z.reshape(-1) for z in atomic_number_lists]).to(self.device, torch.int32).contiguous()
        keys = torch.empty(len(counts), dtype=torch.int64, device=self.device)
        ops.composition_key(Z, off, len(counts), keys)
        return keys

    @staticmethod
    def _cat(a, b):
        if a is None:
            return b
        M = max(a["frac"].shape[1], b["frac"].shape[1])

        def padM(t):
            if t.shape[1] == M:
                return t
            shape = list(t.shape)
            shape[1] = M - t.shape[1]
            return torch.cat([t, torch.zeros(shape, dtype=t.dtype, device=t.device)], dim=1)
        out = {}
        for k in a:
            x, y = a[k], b[k]
            if k in ("frac", "Z"):
                x, y = padM(x), padM(y)
            out[k] = torch.cat([x, y], dim=0)
        return out

    # ------------------------------------------------------------------ reference API
    def extend(self, data, strucs, rewards):
        if len(data) == 0:
            return
        new = self._pack(data, rewards)
        new["key"] = self.keys_of([_atomic_numbers(s) for s in strucs])
        old = None if self._rows is None else {k: v[:self.count] for k, v in self._rows.items()}
        allr = self._cat(old, new)
        n = allr["reward"].shape[0]
        if n > self.MAX_SELECT:
            # beyond the single-CTA kernel's 16 384 rows: the same selection with device sorts (reward desc, stable;
            # first occurrence per key; head; strict cutoff)
            r, key = allr["reward"], allr["key"]
            order = torch.argsort(r, descending=True, stable=True)
            ks, perm = torch.sort(key[order], stable=True)
            first = torch.ones(n, dtype=torch.bool, device=self.device)
            first[1:] = ks[1:] != ks[:-1]
            keep_pos = torch.sort(perm[first]).values[:self.buffer_size]
            sel = order[keep_pos]
            sel = sel[r[sel] > self.reward_cutoff]
            self._rows = {k: v.index_select(0, sel).contiguous() for k, v in allr.items()}
            self.count = int(sel.numel())
            self.M = self._rows["frac"].shape[1]
            return
        idx = torch.empty(n, dtype=torch.int32, device=self.device)
        cnt = torch.zeros(1, dtype=torch.int32, device=self.device)
        ops.replay_select(allr["key"], allr["reward"].contiguous(), n, self.buffer_size, self.reward_cutoff, idx, cnt)
        m = int(cnt)
        sel = idx[:m].long()
        self._rows = {k: v.index_select(0, sel).contiguous() for k, v in allr.items()}
        self.count = m
        self.M = self._rows["frac"].shape[1]

    def sample(self):
        k = min(self.count, self.sample_size)
        if k == 0:
            return [], []
        pick = np.random.choice(self.count, size=k, replace=False)
        r = self._rows
        data = []
        for i in pick.tolist():
            n = int(r["n"][i])
            data.append(CrystalData(r["frac"][i, :n].cpu(), r["Z"][i, :n].cpu().to(torch.int64), r["lengths"][i].view(1, 3).cpu(),
                                    r["angles"][i].view(1, 3).cpu(), torch.tensor(n)))
        return data, r["reward"][torch.as_tensor(pick, device=self.device)].cpu().numpy()       # the stored float64 values

    def memory_purge(self, strucs):
        if self.count == 0 or len(strucs) == 0:
            return
        bad = self.keys_of([_atomic_numbers(s) for s in strucs])
        keep = ~torch.isin(self._rows["key"][:self.count], bad)
        self._rows = {k: v[:self.count][keep].contiguous() for k, v in self._rows.items()}
        self.count = int(keep.sum())

    @property
    def rewards(self):
        return self._rows["reward"][:self.count] if self._rows is not None else torch.zeros(0, device=self.device)

    def __len__(self):
        return self.count
