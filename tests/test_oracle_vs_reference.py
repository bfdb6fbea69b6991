"""CPU, build container only (skipped where /root/reference is absent): the oracle against the UNMODIFIED
reference modules imported under oracle/shims, on fresh seeded inputs (not the committed fixtures)."""
import pytest
import torch

from oracle import diffcsp_oracle as O
from oracle import ref_import as R

pytestmark = pytest.mark.skipif(not R.reference_available(), reason="reference tree not present")


@pytest.mark.parametrize("style", ["fc", "knn"])
def test_forward_matches_live_reference(style):
    hp = O.default_hparams(hidden_dim=64, num_layers=2, num_freqs=8, timesteps=10, edge_style=style, max_neighbors=5)
    torch.manual_seed(0)
    ref = R.build_reference_module(hp)
    sd = O.init_params(hp, seed=3)
    ref.decoder.load_state_dict(sd)
    na = torch.tensor([4, 1, 13, 7])
    n2g = torch.repeat_interleave(torch.arange(4), na)
    g = torch.Generator().manual_seed(1)
    t = O.time_embedding(torch.full((4,), 6), hp["time_dim"])
    a, x = torch.randn(25, 100, generator=g), torch.rand(25, 3, generator=g)
    l = torch.eye(3)[None] * 5 + 0.5 * torch.randn(4, 3, 3, generator=g)
    with torch.no_grad():
        r = ref.decoder(t, a, x, l, na, n2g)
        o = O.cspnet_forward(sd, hp, t, a, x, l, na, n2g)
    for u, v in zip(o, r):
        assert float((u - v).abs().max()) <= 1e-5 * float(v.abs().max())


def test_sampler_matches_live_reference_bit_exact():
    from oracle.make_golden import noise_tape
    hp = O.default_hparams(hidden_dim=64, num_layers=2, num_freqs=8, timesteps=12)
    torch.manual_seed(0)
    ref = R.build_reference_module(hp)
    sd = O.init_params(hp, seed=5)
    ref.decoder.load_state_dict(sd)
    na = [3, 8, 2]
    with noise_tape(9):
        out, _ = ref.sample(R.make_batch(na), step_lr=5e-6)
    sch = O.Schedules(hp, ref.sigma_scheduler.sigmas_norm)
    with torch.no_grad():
        mine = O.sample(sd, hp, sch, na, O.Noise(torch.Generator().manual_seed(9)), step_lr=5e-6)
    for k in ("frac_coords", "lattices", "atom_types"):
        assert torch.equal(mine[k], out[k]), k


def test_long_term_memory_oracle_matches_live_reference():
    """oracle/ltm_oracle.py against the UNMODIFIED memory/ltm.py (pandas) driven with duck-typed structures"""
    import types
    import warnings
    import numpy as np
    from oracle.ltm_oracle import LongTimeMemOracle
    RefLTM = R.import_reference_ltm()

    def struc(formula, elements):
        return types.SimpleNamespace(composition=types.SimpleNamespace(reduced_formula=formula), species=list(elements))

    rng = np.random.default_rng(11)
    names = [("TiO2", ("Ti", "O")), ("FeO", ("Fe", "O")), ("Fe2O3", ("Fe", "O")), ("LiPO4", ("Li", "P", "O")), ("TiO", ("Ti", "O")),
             ("NaCl", ("Na", "Cl"))]
    ref, ora = RefLTM(), LongTimeMemOracle()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")                  # pandas concat-with-empty FutureWarning inside the reference
        for step in range(8):
            pick = [names[i] for i in rng.integers(0, len(names), 30)]
            rew = rng.random(30)
            ref.extend([struc(f, e) for f, e in pick], rew, step)
            ora.extend([f for f, _ in pick], [tuple(sorted(set(e))) for _, e in pick], rew, step)
            q = [names[i] for i in rng.integers(0, len(names), 12)] + [("KBr", ("K", "Br"))]
            qr = rng.random(len(q))
            for method in ("composition", "element_comb"):
                a = ref.div_filter([struc(f, e) for f, e in q], qr, tol=10, buff=20, method=method)
                vals = [f for f, _ in q] if method == "composition" else [tuple(sorted(set(e))) for _, e in q]
                b = ora.div_filter(vals, qr, tol=10, buff=20, method=method)
                assert np.array_equal(a[0], b[0]) and list(a[1]) == b[1] and (a[2], a[3]) == (b[2], b[3])
            for thred, cand in ((0.5, 3), (0.95, 100)):
                m1, m2 = ref.calc_metrics(thred, budget=150, num_candidate=cand), ora.calc_metrics(thred, budget=150, num_candidate=cand)
                assert m1 == m2, (m1, m2)
            assert abs(ref.get_baseline(step) - ora.get_baseline(step)) < 1e-15
            assert len(ref) == len(ora.memory) and len(ref.unique_comps) == len(ora.unique_comps)


def test_replay_buffer_oracle_matches_live_reference():
    """oracle.diffcsp_oracle.ReplayBufferOracle against the UNMODIFIED memory/replay_buffer.py:11-104 (pandas) driven with
    duck-typed structures: extend (dedupe by reduced formula keeping the best, top-K, strict cutoff), sample (numpy's
    global RNG, without replacement), memory_purge — same rows in the same order after every call."""
    import types
    import warnings
    import numpy as np
    RefRB = R.import_reference_replay_buffer()

    def struc(formula, elements):
        return types.SimpleNamespace(composition=types.SimpleNamespace(reduced_formula=formula), species=list(elements))

    names = [("TiO2", ("Ti", "O")), ("FeO", ("Fe", "O")), ("Fe2O3", ("Fe", "O")), ("LiPO4", ("Li", "P", "O")), ("TiO", ("Ti", "O")),
             ("NaCl", ("Na", "Cl")), ("KBr", ("K", "Br")), ("MgO", ("Mg", "O")), ("Al2O3", ("Al", "O")), ("SiC", ("Si", "C"))]
    rng = np.random.default_rng(5)
    for size, ssize, cutoff in ((6, 3, 0.1), (100, 10, 0.1), (4, 8, 0.0)):
        ref, ora = RefRB(size, ssize, cutoff), O.ReplayBufferOracle(size, ssize, cutoff)
        uid = 0
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            for it in range(12):
                pick = [names[i] for i in rng.integers(0, len(names), 7)]
                rew = rng.random(7)
                if it == 3:
                    rew[:] = 0.05          # everything at or under the cutoff
                data = list(range(uid, uid + 7))
                uid += 7
                ref.extend(data, [struc(f, e) for f, e in pick], rew)
                ora.extend(data, [f for f, _ in pick], rew)
                assert ref.buffer["data"].tolist() == [r[0] for r in ora.rows]
                assert ref.buffer["comp"].tolist() == [r[1] for r in ora.rows]
                assert np.array_equal(ref.buffer["reward"].values.astype(float), np.array([r[2] for r in ora.rows]))
                np.random.seed(100 + it)
                d1, r1 = ref.sample()
                np.random.seed(100 + it)
                d2, r2 = ora.sample(np.random)
                assert list(d1) == list(d2) and np.array_equal(np.asarray(r1, dtype=float), np.asarray(r2, dtype=float))
                if it % 4 == 2:
                    bad = [names[i] for i in rng.integers(0, len(names), 2)]
                    ref.memory_purge([struc(f, e) for f, e in bad])
                    ora.memory_purge([f for f, _ in bad])
                    assert ref.buffer["data"].tolist() == [r[0] for r in ora.rows]
                assert len(ref) == len(ora)


def test_reward_scoring_oracle_matches_live_reference(tmp_path):
    """oracle.pipeline_oracle.reward_scoring against the UNMODIFIED rewards/reward.py:33-115 (`Reward.scoring`) driven with
    stub calculators: every target kind, every reduce mode, NaN (failed) samples — bit for bit"""
    import types
    import numpy as np
    from oracle import pipeline_oracle as P
    mod = R.import_reference_reward()
    rng = np.random.default_rng(3)
    B = 40
    raw = [rng.normal(2000, 1500, B), rng.random(B) * 0.4, rng.normal(3.0, 1.5, B)]
    raw[0][[3, 17]] = np.nan
    raw[2][[17, 30]] = np.nan
    cfgs = [dict(name="hhi", target="descending", minv=750, maxv=3250, weight=0.5),
            dict(name="magmom", target="ascending", minv=0.0, maxv=0.25, weight=0.3),
            dict(name="band_gap", target=3.0, minv=0.0, maxv=2.0, weight=0.2)]

    def calc_of(arr):
        return types.SimpleNamespace(calc=lambda samples, label: arr.copy())

    for reduce in ("mean", "min", "weight"):
        for sel in ([0], [0, 1], [0, 1, 2], [2]):
            prop_cfg = [types.SimpleNamespace(calculator=calc_of(raw[i]), **cfgs[i]) for i in sel]
            ref = mod.Reward(root_dir=str(tmp_path), prop_cfg=prop_cfg, reward_threshold=0.8, reduce=reduce)
            r1, d1, f1 = ref.scoring(([None] * B, None), "x")
            r2, d2, f2 = P.reward_scoring([raw[i] for i in sel], [cfgs[i] for i in sel], reduce)
            assert np.array_equal(r1, r2) and np.array_equal(f1, f2), (reduce, sel)
            for k in d1:
                assert np.array_equal(d1[k], d2[k]), k


# ---------------------------------------------------------------- MatterGen adapter (in-tree arithmetic only)
class _FakeBatch:
    """duck-typed ChemGraph batch: fields by key, batch indices per field, batch size"""

    def __init__(self, fields, batch_idx, B):
        self.fields, self._bi, self.B = fields, batch_idx, B

    def __getitem__(self, k):
        return self.fields[k]

    def __contains__(self, k):
        return k in self.fields

    def get_batch_idx(self, k):
        return None if k == "cell" else self._bi

    def get_batch_size(self):
        return self.B


@pytest.mark.skipif(not R.mattergen_adapter_available(), reason="reference MatterGen adapter not present")
def test_mattergen_adapter_oracle_matches_live_reference():
    """oracle/mattergen_oracle.py against the UNMODIFIED models/mattergen/{pl_module,loss}.py (stub mattergen leaves):
    the fine-tune time of every index, the per-sample weighted loss aggregation, the KL proxy — bit for bit"""
    import types
    from oracle import mattergen_oracle as MO
    plm, loss_mod = R.import_reference_mattergen_adapter()
    g = torch.Generator().manual_seed(0)
    B, na = 5, torch.tensor([3, 1, 7, 20, 4])
    N = int(na.sum())
    bi = torch.repeat_interleave(torch.arange(B), na)
    seen = {}

    class Corr:
        T = 1.0
        corruptions = {"pos": "cp", "cell": "cc", "atomic_numbers": "ca"}

        def sample_marginal(self, batch, t):
            seen["t"] = t.clone()
            return "noisy"

    dm = types.SimpleNamespace(pre_corruption_fn=lambda b: b, corruption=Corr(), _get_device=lambda b: torch.device("cpu"),
                               model=None, loss_fn=None)
    ref = plm.MatterGenModule(diffusion_module=dm)
    batch = _FakeBatch({}, bi, B)
    for ts in (0, 1, 17, 499, 500, 998, 999):
        noisy, b2, t = ref.add_noise(batch, ts)
        assert noisy == "noisy" and b2 is batch and t.shape == (B,)
        assert torch.equal(t, torch.full((B,), float(MO.finetune_time(1.0, ts)))), ts
    # per-sample aggregation with stub per-field losses (what MaterialsLoss would compute is un-vendored)
    vals = {"pos": torch.rand(B, generator=g), "cell": torch.rand(B, generator=g), "atomic_numbers": torch.rand(B, generator=g)}
    sl = loss_mod.SampleLoss()
    sl.loss_fns = {k: (lambda k_: (lambda **kw: vals[k_]))(k) for k in sl.loss_fns}
    fb = _FakeBatch({k: None for k in vals}, bi, B)
    agg, metrics = sl(multi_corruption=Corr(), batch=fb, noisy_batch=fb, score_model_output=fb, t=torch.zeros(B))
    agg2, metrics2 = MO.aggregate_sample_loss({k: vals[k] for k in sl.loss_fns}, sl.loss_weights)
    assert torch.equal(agg, agg2) and sl.loss_weights == MO.DEFAULT_WEIGHTS
    assert all(torch.equal(metrics[k], metrics2[k]) for k in metrics)
    # KL proxy
    ap = {"pos": torch.randn(N, 3, generator=g), "cell": torch.randn(B, 3, 3, generator=g), "atomic_numbers": torch.randn(N, 101, generator=g)}
    pp = {k: v + 0.1 * torch.randn(v.shape, generator=g) for k, v in ap.items()}
    kl = ref.calc_kl_reg(ap, pp, batch)
    assert torch.equal(kl, MO.kl_reg(ap, pp, bi, B))
