"""GPU: the post-sampling rows of SURVEY.md §8(f) — validity pre-filter, composition rewards, vectorised replay packing,
extxyz dumps — against oracle/pipeline_oracle.py (numpy float64), all through the C ABI."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _crystals(n_items, seed, tiny_cells=False):
    from matinvent_b200.models.diffcsp.sample import CrystalData
    g = torch.Generator().manual_seed(seed)
    out = []
    for i in range(n_items):
        n = int(torch.randint(1, 21, (1,), generator=g))
        lengths = 2.0 + 9.0 * torch.rand(1, 3, generator=g)
        if tiny_cells and i % 5 == 0:
            lengths = lengths * 0.08                      # cells shorter than the 0.5 A cutoff: self-image distances
        if i % 7 == 3:
            lengths[0, int(torch.randint(0, 3, (1,), generator=g))] = 24.0 + 3.0 * float(torch.rand(1, generator=g))   # around 25 A
        if i % 11 == 5:
            lengths[0, 1] = 41.0
        angles = 60.0 + 60.0 * torch.rand(1, 3, generator=g)
        out.append(CrystalData(torch.rand(n, 3, generator=g), torch.randint(1, 101, (n,), generator=g), lengths, angles,
                               torch.tensor(n)))
    return out


def test_validity_prefilter_matches_oracle():
    from oracle import pipeline_oracle as P
    from matinvent_b200.pipeline.filters import invalid_filter, validity_masks
    data = _crystals(300, 1, tiny_cells=True)
    # exact edge of the in-tree rule
    data[0].lengths = torch.tensor([[25.0, 3.0, 3.0]])
    data[1].lengths = torch.tensor([[24.999998, 3.0, 3.0]])
    cell_ok, struc_ok, dmin = validity_masks(data, device="cuda")
    near = 0
    for i, d in enumerate(data):
        assert bool(cell_ok[i]) == P.cell_length_ok(d.lengths.reshape(-1).tolist()), i
        L = P.lattice_matrix(d.lengths.reshape(-1).tolist(), d.angles.reshape(-1).tolist())
        ref_d = P.min_periodic_distance(d.frac_coords.numpy(), L)
        assert abs(dmin[i] - ref_d) <= 2e-5 * max(1.0, ref_d), (i, dmin[i], ref_d)
        if abs(ref_d - 0.5) < 1e-4 or abs(abs(np.linalg.det(L)) - 0.1) < 1e-4:
            near += 1
            continue                                          # fp32 vs fp64 on the threshold itself
        assert bool(struc_ok[i]) == P.structure_validity(d.frac_coords.numpy(), d.lengths.reshape(-1).tolist(),
                                                         d.angles.reshape(-1).tolist()), i
    assert near < 5 and cell_ok.sum() not in (0, len(data)) and struc_ok.sum() not in (0, len(data))
    assert not cell_ok[0] and cell_ok[1]
    kept, kept_s = invalid_filter(data, data, device="cuda")
    assert len(kept) == int((cell_ok & struc_ok).sum()) and all(a is b for a, b in zip(kept, kept_s))
    mask = invalid_filter(data, data, return_mask=True, structure_validity=False, smact_validity=lambda s: int(s.num_atoms) % 2 == 0,
                          device="cuda")
    assert np.array_equal(mask, cell_ok & np.array([int(d.num_atoms) % 2 == 0 for d in data]))
    assert invalid_filter([], [], device="cuda") == ([], [])


@pytest.mark.parametrize("reduce", ["mean", "min", "weight"])
def test_composition_reward_matches_oracle(reduce):
    from oracle import pipeline_oracle as P
    from matinvent_b200.rewards import CompositionReward, synthetic_table
    from matinvent_b200.rewards.elements import ATOMIC_MASS
    data = _crystals(500, 2)
    cfgs = [dict(name="hhi", table=synthetic_table("hhi"), weights="mass", target="descending", minv=750, maxv=3250, weight=0.6),
            dict(name="magmom", table=synthetic_table("magmom"), weights="atom", target="ascending", minv=0.0, maxv=0.25, weight=0.3),
            dict(name="hhi_t", table=synthetic_table("hhi"), weights="mass", target=2000.0, minv=0.0, maxv=1500.0, weight=0.1)]
    rw = CompositionReward(prop_cfg=cfgs, reward_threshold=0.8, reduce=reduce, device="cuda")
    r, props, failed = rw.scoring((data, None), "t")
    raw = [[P.composition_property(d.atom_types.tolist(), c["table"], ATOMIC_MASS, c["weights"]) for d in data] for c in cfgs]
    r2, props2, failed2 = P.reward_scoring(raw, cfgs, reduce)
    assert np.array_equal(failed, failed2) and 0 < failed.sum() < len(data)        # Z > 94 has no table entry
    for k in props2:
        assert np.allclose(props[k], props2[k], rtol=1e-13, atol=0), k
    assert np.allclose(r, r2, rtol=1e-12, atol=1e-15)
    assert (r[failed] == 0).all() and rw.threshold == 0.8
    r0, p0, f0 = rw.scoring(([], None))
    assert len(r0) == 0 and len(f0) == 0


def test_replay_pack_vectorised_and_float64_rewards():
    """10 k crystals go into the buffer with one H2D copy per field; rewards keep their float64 values (near-ties and the
    strict cutoff are decided in double like the pandas column of memory/replay_buffer.py:60-71)"""
    from matinvent_b200.memory import ReplayBuffer
    data = _crystals(10000, 3)
    rng = np.random.default_rng(0)
    rewards = rng.random(10000)
    rewards[:3] = [0.5 + 1e-12, 0.5, 0.5 - 1e-12]                      # equal in float32
    buf = ReplayBuffer(buffer_size=100, sample_size=10, reward_cutoff=0.5, device="cuda")
    for d in data[:3]:
        d.atom_types = torch.tensor([8, 8, 22][: int(d.num_atoms)] + [22] * max(0, int(d.num_atoms) - 3))
    data[1].atom_types = torch.tensor([3])
    data[1].num_atoms, data[1].frac_coords = torch.tensor(1), data[1].frac_coords[:1]
    data[2].atom_types = torch.tensor([4])
    data[2].num_atoms, data[2].frac_coords = torch.tensor(1), data[2].frac_coords[:1]
    buf.extend(data[:3], data[:3], rewards[:3])
    assert len(buf) == 1 and float(buf.rewards[0]) == 0.5 + 1e-12          # strict cutoff in double: only the first survives
    buf = ReplayBuffer(buffer_size=100, sample_size=10, reward_cutoff=0.1, device="cuda")
    buf.extend(data, data, rewards)
    from oracle.diffcsp_oracle import ReplayBufferOracle, reduced_composition_key
    orc = ReplayBufferOracle(100, 10, 0.1)
    orc.extend(list(range(10000)), [reduced_composition_key(d.atom_types.tolist()) for d in data], rewards)
    assert len(buf) == len(orc) == 100
    assert np.array_equal(buf.rewards.cpu().numpy(), np.array([r[2] for r in orc.rows]))
    for row, (i, _, _) in enumerate(orc.rows[:20]):
        n = int(buf._rows["n"][row])
        assert n == int(data[i].num_atoms)
        assert torch.equal(buf._rows["Z"][row, :n].cpu().long(), data[i].atom_types.long())
        assert torch.equal(buf._rows["frac"][row, :n].cpu(), data[i].frac_coords.float())
    np.random.seed(0)
    picked, rew = buf.sample()
    assert len(picked) == 10 and rew.dtype == np.float64 and set(rew.tolist()) <= set(r[2] for r in orc.rows)
    # the > 16 384 row path (device sorts) selects the same rows as the single-CTA kernel
    more = _crystals(9000, 4)
    r_more = rng.random(9000)
    buf.extend(more, more, r_more)
    orc.extend(list(range(10000, 19000)), [reduced_composition_key(d.atom_types.tolist()) for d in more], r_more)
    assert np.array_equal(buf.rewards.cpu().numpy(), np.array([r[2] for r in orc.rows]))


def test_extxyz_dump_of_sampled_crystals(tmp_path):
    from oracle import pipeline_oracle as P
    from matinvent_b200.pipeline.utils import read_extxyz, save_structures
    from matinvent_b200.rewards.elements import SYMBOLS
    data = _crystals(12, 5)
    path = save_structures(data, str(tmp_path), "step_0000_valid.extxyz")
    frames = read_extxyz(path)
    assert len(frames) == 12
    for d, (sym, pos, cell) in zip(data, frames):
        L = P.lattice_matrix(d.lengths.reshape(-1).tolist(), d.angles.reshape(-1).tolist())
        assert np.allclose(np.array(cell), L, rtol=1e-12, atol=1e-12)
        assert sym == [SYMBOLS[int(z)] for z in d.atom_types.tolist()]
        assert np.allclose(np.array(pos), d.frac_coords.double().numpy() @ L, atol=2e-8)
