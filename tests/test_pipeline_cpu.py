"""CPU: host logic of the pipeline rows — crystal partitioning across ranks, the extxyz writer, the pipeline oracle's own
known answers.  No GPU, no compute calls into the library."""
import numpy as np
import torch


def test_partition_crystals_never_starves_a_rank():
    """one 20-atom crystal crosses several total*r/world thresholds at once; every rank must still get a
    crystal when there are enough, slices stay contiguous and cover the batch, and stay balanced by sum n^2"""
    from matinvent_b200.models.diffcsp.finetune import partition_crystals
    from matinvent_b200.models.diffcsp.sample import ATOM_DIST
    rs = np.random.RandomState(0)
    for _ in range(2000):
        B = int(rs.randint(1, 40))
        na = rs.choice(21, B, p=ATOM_DIST["mp_20"]).tolist()
        for world in (2, 4, 8):
            parts = partition_crystals(na, world)
            assert len(parts) == world and parts[0][0] == 0 and parts[-1][1] == B
            assert all(parts[k][1] == parts[k + 1][0] for k in range(world - 1))
            if B >= world:
                assert all(hi > lo for lo, hi in parts), (na, world, parts)
            else:
                assert all(hi >= lo for lo, hi in parts) and sum(hi > lo for lo, hi in parts) == B
    assert partition_crystals([1, 10], 2) == [(0, 1), (1, 2)]
    na = rs.choice(21, 2048, p=ATOM_DIST["mp_20"]).tolist()
    for world in (2, 4, 8):
        e = [sum(n * n for n in na[lo:hi]) for lo, hi in partition_crystals(na, world)]
        assert max(e) / (sum(e) / world) < 1.02


def test_extxyz_writer_layout(tmp_path):
    from matinvent_b200.models.diffcsp.sample import CrystalData
    from matinvent_b200.pipeline.utils import read_extxyz, save_structures
    d = CrystalData(torch.tensor([[0.0, 0.0, 0.0], [0.5, 0.5, 0.5]]), torch.tensor([11, 17]), torch.tensor([[4.0, 4.0, 4.0]]),
                    torch.tensor([[90.0, 90.0, 90.0]]), torch.tensor(2))
    path = save_structures([d, d], str(tmp_path), "a.extxyz")
    lines = open(path).read().splitlines()
    assert lines[0] == "2" and lines[1].startswith('Lattice="4.0 0.0 ') and 'Properties=species:S:1:pos:R:3 pbc="T T T"' in lines[1]
    assert lines[2].split()[0] == "Na" and lines[3].split()[0] == "Cl" and len(lines) == 8
    (sym, pos, cell), _ = read_extxyz(path)
    assert sym == ["Na", "Cl"] and np.allclose(pos[1], [2.0, 2.0, 2.0], atol=1e-6) and np.allclose(np.diag(cell), 4.0)


def test_pipeline_oracle_known_answers():
    from oracle import pipeline_oracle as P
    # rock salt, a = 5.64: nearest neighbour a/2; primitive cubic with one atom: its own image at a
    fcc = [[0, 0, 0], [0.5, 0.5, 0], [0.5, 0, 0.5], [0, 0.5, 0.5], [0.5, 0, 0], [0, 0.5, 0], [0, 0, 0.5], [0.5, 0.5, 0.5]]
    L = P.lattice_matrix([5.64] * 3, [90.0] * 3)
    assert abs(P.min_periodic_distance(fcc, L) - 2.82) < 1e-12
    assert abs(P.min_periodic_distance([[0.3, 0.2, 0.9]], P.lattice_matrix([0.4, 3, 3], [90.0] * 3)) - 0.4) < 1e-12
    assert not P.structure_validity([[0.3, 0.2, 0.9]], [0.4, 3, 3], [90.0] * 3) and P.structure_validity(fcc, [5.64] * 3, [90.0] * 3)
    assert P.cell_length_ok([24.99, 3, 3]) and not P.cell_length_ok([25.0, 3, 3])
    # rewards/reward.py docstring-level cases: descending 750..3250, value 2000 -> 0.5; NaN -> failed, reward 0
    r, d, f = P.reward_scoring([np.array([2000.0, np.nan, 100.0])], [dict(name="hhi", target="descending", minv=750, maxv=3250)])
    assert r.tolist() == [0.5, 0.0, 1.0] and f.tolist() == [False, True, False] and d["hhi"].tolist() == [2000.0, 0.0, 100.0]
    # SiO2: mass fractions 28.085 / 60.083 and 31.998 / 60.083
    mass = np.zeros(101); mass[14], mass[8] = 28.085, 15.999
    tab = np.zeros(101); tab[14], tab[8] = 1000.0, 500.0
    v = P.composition_property([14, 8, 8], tab, mass, "mass")
    assert abs(v - (28.085 * 1000 + 31.998 * 500) / 60.083) < 1e-9
    assert abs(P.composition_property([14, 8, 8], tab, mass, "atom") - (1000 + 2 * 500) / 3) < 1e-9


def test_pack_unpack_crystals_roundtrip():
    """the inter-rank transport of sampled crystals (MatInvent.sample_step): five concatenated tensors per shard"""
    import torch
    from matinvent_b200.models.diffcsp.sample import CrystalData, pack_crystals, unpack_crystals
    g = torch.Generator().manual_seed(0)
    data = [CrystalData(torch.rand(n, 3, generator=g), torch.randint(1, 101, (n,), generator=g), torch.rand(1, 3, generator=g),
                        torch.rand(1, 3, generator=g), torch.tensor(n)) for n in (3, 1, 20, 7)]
    back = unpack_crystals(pack_crystals(data))
    assert len(back) == len(data)
    for a, b in zip(data, back):
        assert torch.equal(a.frac_coords, b.frac_coords) and torch.equal(a.atom_types, b.atom_types)
        assert torch.equal(a.lengths, b.lengths) and torch.equal(a.angles, b.angles) and int(a.num_atoms) == int(b.num_atoms)
    assert unpack_crystals(pack_crystals([])) == []
