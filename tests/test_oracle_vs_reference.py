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
