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
