"""CPU: pins the oracle (oracle/diffcsp_oracle.py) to golden vectors produced by the UNMODIFIED reference
(tests/golden/*.pt, generator: oracle/make_golden.py) and to the only known answers the reference tree itself
holds: the `repeat_blocks` docstring examples (models/diffcsp/utils.py:208-226)."""
import torch

from oracle import diffcsp_oracle as O
from oracle.ref_import import make_batch


def test_repeat_blocks_docstring_examples():
    rb = lambda *a, **k: O.repeat_blocks(*a, **k).tolist()
    assert rb([1, 3, 2], [3, 2, 3], continuous_indexing=False) == [0, 0, 0, 0, 1, 2, 0, 1, 2, 0, 1, 0, 1, 0, 1]
    assert rb([1, 3, 2], [3, 2, 3], continuous_indexing=True) == [0, 0, 0, 1, 2, 3, 1, 2, 3, 4, 5, 4, 5, 4, 5]
    assert rb([1, 3, 2], [3, 2, 3], continuous_indexing=True, repeat_inc=4) == [0, 4, 8, 1, 2, 3, 5, 6, 7, 4, 5, 8, 9, 12, 13]
    assert rb([1, 3, 2], [3, 2, 3], continuous_indexing=True, start_idx=5) == [5, 5, 5, 6, 7, 8, 6, 7, 8, 9, 10, 9, 10, 9, 10]
    assert rb([1, 3, 2], [3, 2, 3], continuous_indexing=True, block_inc=1) == [0, 0, 0, 2, 3, 4, 2, 3, 4, 6, 7, 6, 7, 6, 7]
    assert rb([0, 3, 2], [3, 2, 3], continuous_indexing=True) == [0, 1, 2, 0, 1, 2, 3, 4, 3, 4, 3, 4]
    assert rb([2, 3, 2], [2, 0, 2], continuous_indexing=True) == [0, 1, 0, 1, 5, 6, 5, 6]


def test_schedules(gold_small):
    gs = gold_small
    hp = gs["hp"]
    b = O.beta_tables(hp["timesteps"], hp["beta_mode"])
    for k, v in gs["beta"].items():
        assert torch.equal(b[k], v), k
    s = O.sigma_tables(hp["timesteps"], hp["sigma_begin"], hp["sigma_end"], gs["sigmas_norm"])
    assert torch.equal(s["sigmas"], gs["sigma_sigmas"])


def _fwd(gs, hp, case):
    na = case["num_atoms"]
    n2g = torch.repeat_interleave(torch.arange(len(na)), na)
    with torch.no_grad():
        return O.cspnet_forward(gs["sd"], hp, case["temb"], case["a"], case["x"], case["l"], na, n2g), n2g


def test_forward_fc_bit_exact(gold_small):
    gs = gold_small
    c = gs["forward_fc"]
    (pl, px, pt), n2g = _fwd(gs, gs["hp"], c)
    assert torch.equal(pl, c["ref_pred_l"]) and torch.equal(px, c["ref_pred_x"]) and torch.equal(pt, c["ref_pred_t"])
    e, fd = O.gen_edges(gs["hp"], c["num_atoms"], c["x"], c["l"], n2g)
    assert torch.equal(e.int(), c["ref_edges"]) and torch.equal(fd, c["ref_frac_diff"])


def test_forward_baseline_batch(gold_baseline):
    """the oracle against the unmodified reference on the benchmark batch (256 mp_20 crystals, 34 445 edges, full-size
    net regenerated from its seed and pinned by checksums)"""
    from conftest import baseline_inputs
    gb = gold_baseline
    sd = O.init_params(gb["hp"], gb["seed_weights"])
    for k, v in sd.items():
        assert abs(float(v.double().abs().sum()) - gb["checksums"][k]) <= 1e-9 * max(1.0, gb["checksums"][k]), k
    na, t, a, x, l, n2g = baseline_inputs(gb)
    assert int((na * na).sum()) == gb["edges"] == 34445
    with torch.no_grad():
        pl, px, pt = O.cspnet_forward(sd, gb["hp"], t, a, x, l, na, n2g)
    for mine, ref in ((pl, gb["ref_pred_l"]), (px, gb["ref_pred_x"]), (pt[::8], gb["ref_pred_t_rows8"])):
        assert float((mine - ref).abs().max()) <= 2e-6 * float(ref.abs().max())


def test_forward_knn(gold_small):
    gs = gold_small
    c = gs["forward_knn"]
    (pl, px, pt), n2g = _fwd(gs, gs["hp_knn"], c)
    e, fd = O.gen_edges(gs["hp_knn"], c["num_atoms"], c["x"], c["l"], n2g)
    assert torch.equal(e.int(), c["ref_edges"]) and torch.allclose(fd, c["ref_frac_diff"], atol=1e-6)
    for a, b in ((pl, c["ref_pred_l"]), (px, c["ref_pred_x"]), (pt, c["ref_pred_t"])):
        assert float((a - b).abs().max()) <= 1e-5 * float(b.abs().max())


def test_radius_graph_pbc(gold_rg):
    rg = gold_rg
    for K in (4, 20):
        ei, uc, nb = O.radius_graph_pbc(rg["cart"], rg["lattices"], rg["num_atoms"], K)
        ref = rg["ref_K%d" % K]
        assert torch.equal(ei.int(), ref["edge_index"]) and torch.equal(uc.to(torch.int8), ref["cell"])
        assert torch.equal(nb, ref["per_image"])


def test_ft_timestep_loss_and_grads(gold_small):
    gs = gold_small
    hp, ft = gs["hp"], gs["ft"]
    sch = O.Schedules(hp, gs["sigmas_norm"])
    batch = make_batch(ft["num_atoms"].tolist(), **ft["crystals"])
    sd = {k: v.clone().requires_grad_(True) for k, v in gs["sd"].items()}
    noise = O.Noise(torch.Generator().manual_seed(ft["noise_seed"]))
    loss, parts = O.ft_timestep_loss(sd, gs["sd_prior"], hp, sch, batch, batch.reward, ft["t_idx"], noise,
                                     ft["sigma"], ft["accum"])
    loss.backward()
    (temb, a_t, x_t, l_t, _, _), (rl, tx, rt), _ = parts["noised"]
    assert torch.equal(temb, ft["ref_temb"]) and torch.equal(a_t, ft["ref_a_t"]) and torch.equal(x_t, ft["ref_x_t"])
    assert torch.equal(l_t, ft["ref_l_t"]) and torch.equal(tx, ft["ref_tar_x"])
    assert torch.allclose(parts["sample_loss"], ft["ref_sample_loss"], rtol=1e-6, atol=0)
    assert torch.allclose(parts["kl"], ft["ref_kl"], rtol=1e-6, atol=0)
    assert abs(float(loss.detach()) - float(ft["ref_loss"])) <= 1e-6 * abs(float(ft["ref_loss"]))
    for k, v in gs["ft_grads"].items():
        g = sd[k].grad
        assert float((g - v).abs().max()) <= 1e-5 * float(v.abs().max()) + 1e-12, k


def test_sample_T40_shared_tape(gold_small):
    gs = gold_small
    hp, s = gs["hp"], gs["sample"]
    sch = O.Schedules(hp, gs["sigmas_norm"])
    noise = O.Noise(torch.Generator().manual_seed(s["seed"]))
    with torch.no_grad():
        out, traj = O.sample(gs["sd"], hp, sch, s["num_atoms"].tolist(), noise, step_lr=s["step_lr"], return_traj=True)
    T = hp["timesteps"]
    for t, ref in s["ref_traj"].items():
        if t == T:
            continue
        mine = traj[T - 1 - t]
        assert torch.equal(mine["frac_coords"], ref["frac_coords"]) and torch.equal(mine["lattices"], ref["lattices"]), t
    assert torch.equal(out["frac_coords"], s["ref_frac_coords"])
    assert torch.equal(out["lattices"], s["ref_lattices"]) and torch.equal(out["atom_types"], s["ref_atom_types"])


def test_postprocess_and_composition_key():
    out = dict(frac_coords=torch.rand(5, 3), lattices=torch.eye(3)[None].repeat(2, 1, 1) * 4.0,
               atom_types=torch.randn(5, 100), num_atoms=torch.tensor([2, 3]))
    cr = O.generate_postprocess(out)
    assert len(cr) == 2 and cr[1]["atom_types"].shape == (3,) and torch.allclose(cr[0]["angles"], torch.full((1, 3), 90.0))
    assert O.reduced_composition_key([26, 26, 8, 8, 8]) == O.reduced_composition_key([8] * 6 + [26] * 4)
    assert O.reduced_composition_key([26, 26, 8, 8, 8]) != O.reduced_composition_key([26] * 3 + [8] * 4)
