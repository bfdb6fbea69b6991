"""GPU parity: the CUDA path against golden vectors produced by the UNMODIFIED reference
(tests/golden/*.pt, made by oracle/make_golden.py) and against the oracle restatement.

Tolerances (BASELINE.json north_star): fractional coordinates 1e-4 (wrapped distance), lattices 1e-4 rel
(max-norm), atom-type indices bit-exact; per-crystal losses and gradients 1e-4 rel."""
import pytest
import torch

from conftest import build_module, rel_err, wrapped_err

pytestmark = pytest.mark.gpu


def _types(a):
    return torch.argmax(a.detach().cpu(), dim=-1) + 1


def _forward_case(m, case):
    na = case["num_atoms"]
    n2g = torch.repeat_interleave(torch.arange(len(na)), na)
    with torch.no_grad():
        return m.decoder(case["temb"].cuda(), case["a"].cuda(), case["x"].cuda(), case["l"].cuda(), na, n2g)


def test_forward_small_fc(gold_small):
    gs = gold_small
    m = build_module(gs["hp"], gs["sd"], gs["sigmas_norm"])
    c = gs["forward_fc"]
    pl, px, pt = _forward_case(m, c)
    assert rel_err(pl, c["ref_pred_l"]) < 2e-5
    assert rel_err(px, c["ref_pred_x"]) < 2e-5
    assert rel_err(pt, c["ref_pred_t"]) < 2e-5


def test_state_dict_roundtrip(gold_small):
    gs = gold_small
    m = build_module(gs["hp"], gs["sd"], gs["sigmas_norm"])
    sd = m.decoder.state_dict()
    assert set(sd.keys()) == set(gs["sd"].keys())
    for k, v in gs["sd"].items():
        assert torch.equal(sd[k].cpu(), v), k
    full = m.state_dict()
    assert "decoder.csp_layer_0.edge_mlp.0.weight" in full and "sigma_scheduler.sigmas_norm" in full


def test_schedules_match_reference_buffers(gold_small):
    gs = gold_small
    m = build_module(gs["hp"], gs["sd"], gs["sigmas_norm"])
    for k, v in gs["beta"].items():
        assert torch.equal(getattr(m.beta_scheduler, k).cpu(), v), k
    assert torch.equal(m.sigma_scheduler.sigmas.cpu(), gs["sigma_sigmas"])


@pytest.mark.parametrize("noise_graph", [True, False])
def test_sample_small_T40(gold_small, noise_graph):
    from matinvent_b200.models.diffcsp import TapeNoise
    from oracle.ref_import import make_batch
    gs = gold_small
    s = gs["sample"]
    m = build_module(gs["hp"], gs["sd"], gs["sigmas_norm"])
    out, traj = m.sample(make_batch(s["num_atoms"].tolist()), step_lr=s["step_lr"],
                         noise=TapeNoise("cuda", seed=s["seed"]), use_cuda_graph=noise_graph, return_traj=True)
    for t, ref in s["ref_traj"].items():
        assert wrapped_err(traj[t]["frac_coords"], ref["frac_coords"]) < 1e-4, t
        assert rel_err(traj[t]["lattices"], ref["lattices"]) < 1e-4, t
    assert wrapped_err(out["frac_coords"], s["ref_frac_coords"]) < 1e-4
    assert rel_err(out["lattices"], s["ref_lattices"]) < 1e-4
    assert torch.equal(_types(out["atom_types"]), _types(s["ref_atom_types"]))


def test_sample_graph_replay_equals_eager(gold_small):
    from matinvent_b200.models.diffcsp import TapeNoise
    from oracle.ref_import import make_batch
    gs = gold_small
    m = build_module(gs["hp"], gs["sd"], gs["sigmas_norm"])
    b = make_batch([5, 20, 2, 11])
    o1, _ = m.sample(b, step_lr=5e-6, noise=TapeNoise("cuda", seed=3), use_cuda_graph=True)
    o2, _ = m.sample(b, step_lr=5e-6, noise=TapeNoise("cuda", seed=3), use_cuda_graph=False)
    for k in ("frac_coords", "lattices", "atom_types"):
        assert torch.equal(o1[k], o2[k]), k


def test_ft_timestep_gradients_small(gold_small):
    """One inner iteration of MatInvent.ft_step (pipeline/mat_invent.py:152-164) through the plugin API with
    torch autograd on top of the hand-written backward; every parameter gradient vs the reference's."""
    from matinvent_b200.models.diffcsp import TapeNoise
    from oracle.ref_import import make_batch
    gs = gold_small
    ft = gs["ft"]
    agent = build_module(gs["hp"], gs["sd"], gs["sigmas_norm"])
    prior = build_module(gs["hp"], gs["sd_prior"], gs["sigmas_norm"])
    for p in prior.parameters():
        p.requires_grad = False
    batch = make_batch(ft["num_atoms"].tolist(), **ft["crystals"])
    batch.reward = batch.reward.cuda()
    noised = agent.add_noise(batch, ft["t_idx"], noise=TapeNoise("cuda", seed=ft["noise_seed"]))
    sample_loss, agent_pred = agent.calc_sample_loss(noised)
    _, prior_pred = prior.calc_sample_loss(noised)
    kl = agent.calc_kl_reg(agent_pred, prior_pred, batch)
    loss = (batch.reward * sample_loss + kl * (1.1 - batch.reward) * ft["sigma"]).mean() / ft["accum"]
    loss.backward()
    assert rel_err(sample_loss, ft["ref_sample_loss"]) < 1e-4
    assert rel_err(kl, ft["ref_kl"]) < 1e-4
    assert abs(float(loss) - float(ft["ref_loss"])) < 1e-5 * abs(float(ft["ref_loss"]))
    for k in range(3):
        assert rel_err(agent_pred[k], ft["ref_agent_pred"][k]) < 2e-5
        assert rel_err(prior_pred[k], ft["ref_prior_pred"][k]) < 2e-5
    grads = agent.decoder.reference_named_grads()
    assert set(grads.keys()) == set(gs["ft_grads"].keys())
    worst = max((rel_err(grads[k], v), k) for k, v in gs["ft_grads"].items())
    assert worst[0] < 1e-4, worst
    # second backward accumulates (grad accumulation over timesteps, mat_invent.py:163-167)
    noised = agent.add_noise(batch, ft["t_idx"], noise=TapeNoise("cuda", seed=ft["noise_seed"]))
    sl2, ap2 = agent.calc_sample_loss(noised)
    kl2 = agent.calc_kl_reg(ap2, prior_pred, batch)
    ((batch.reward * sl2 + kl2 * (1.1 - batch.reward) * ft["sigma"]).mean() / ft["accum"]).backward()
    g2 = agent.decoder.reference_named_grads()
    worst = max((rel_err(g2[k], 2 * v), k) for k, v in gs["ft_grads"].items())
    assert worst[0] < 1e-4, worst


def _check_grad_projections(grads, proj, tol=1e-4):
    """seeded random projections <g, r> of every gradient tensor against the unmodified reference's (oracle/make_golden.py
    grad_projections): |<g - g_ref, r>| <= tol * ||g_ref||_2, i.e. a relative-L2 bar that sees sign flips, transposes and
    permuted blocks anywhere in the tensor"""
    import zlib
    worst = (0.0, None)
    for k, chk in proj.items():
        v = grads[k].detach().double().cpu().reshape(-1)
        for j, ref in enumerate(chk["proj"]):
            g = torch.Generator().manual_seed(zlib.crc32(("%s#%d" % (k, j)).encode()))
            r = torch.randn(v.numel(), generator=g, dtype=torch.float64)
            err = abs(float(torch.dot(v, r)) - ref) / (chk["norm"] + 1e-300)
            worst = max(worst, (err, k))
            assert err < tol, (k, j, err)
        assert abs(float(v.norm()) - chk["norm"]) < tol * chk["norm"] + 1e-30, k
    return worst


def _full_module(gold_full, which=0):
    from oracle import diffcsp_oracle as O
    hp = gold_full["hp"]
    sd = O.init_params(hp, gold_full["seeds"][which])
    cs = gold_full["checksums" if which == 0 else "checksums_prior"]
    for k, v in sd.items():
        assert abs(float(v.double().abs().sum()) - cs[k]) <= 1e-9 * max(1.0, cs[k]), "weight regeneration drifted: " + k
    sn = torch.load(__import__("os").path.join(__import__("conftest").GOLD, "sigmas_norm_T1000.pt"))["sigmas_norm"]
    return build_module(hp, sd, sn)


def test_forward_full_size(gold_full):
    m = _full_module(gold_full)
    c = gold_full["forward_fc"]
    pl, px, pt = _forward_case(m, c)
    assert rel_err(pl, c["ref_pred_l"]) < 2e-5
    assert rel_err(px, c["ref_pred_x"]) < 2e-5
    assert rel_err(pt, c["ref_pred_t"]) < 2e-5


def test_ft_gradients_full_size(gold_full):
    from matinvent_b200.models.diffcsp import TapeNoise
    from oracle.ref_import import make_batch
    ft = gold_full["ft"]
    agent, prior = _full_module(gold_full, 0), _full_module(gold_full, 1)
    for p in prior.parameters():
        p.requires_grad = False
    batch = make_batch(ft["num_atoms"].tolist(), **ft["crystals"])
    batch.reward = batch.reward.cuda()
    noised = agent.add_noise(batch, ft["t_idx"], noise=TapeNoise("cuda", seed=ft["noise_seed"]))
    sample_loss, agent_pred = agent.calc_sample_loss(noised)
    _, prior_pred = prior.calc_sample_loss(noised)
    kl = agent.calc_kl_reg(agent_pred, prior_pred, batch)
    ((batch.reward * sample_loss + kl * (1.1 - batch.reward) * ft["sigma"]).mean() / ft["accum"]).backward()
    assert rel_err(sample_loss, ft["ref_sample_loss"]) < 1e-4 and rel_err(kl, ft["ref_kl"]) < 1e-4
    grads = agent.decoder.reference_named_grads()
    for k, chk in gold_full["ft_grad_checks"].items():
        g = grads[k].cpu()
        assert abs(float(g.double().abs().sum()) - chk["abs_sum"]) < 1e-4 * chk["abs_sum"] + 1e-12, k
        scale = float(g.abs().max()) + 1e-30
        assert float((g.reshape(-1)[:64] - chk["head"]).abs().max()) < 1e-4 * scale, k
    worst = _check_grad_projections(grads, gold_full["ft_grad_proj"])
    print("full-size fine-tune gradients vs reference: worst projection error %.2e of the tensor's norm (%s)" % worst)


def test_sample_full_size_1000_steps(gold_full):
    """The headline parity case: full-size net, 1000 reverse steps, shared noise tape."""
    from matinvent_b200.models.diffcsp import TapeNoise
    from oracle.ref_import import make_batch
    s = gold_full["sample_T1000"]
    m = _full_module(gold_full)
    out, traj = m.sample(make_batch(s["num_atoms"].tolist()), step_lr=s["step_lr"],
                         noise=TapeNoise("cuda", seed=s["seed"]), return_traj=True)
    for t, ref in s["ref_traj"].items():
        assert wrapped_err(traj[t]["frac_coords"], ref["frac_coords"]) < 1e-4, t
        assert rel_err(traj[t]["lattices"], ref["lattices"]) < 1e-4, t
    print("1000-step parity: frac %.2e  lattice %.2e" % (wrapped_err(out["frac_coords"], s["ref_frac_coords"]),
                                                         rel_err(out["lattices"], s["ref_lattices"])))
    assert wrapped_err(out["frac_coords"], s["ref_frac_coords"]) < 1e-4
    assert rel_err(out["lattices"], s["ref_lattices"]) < 1e-4
    assert torch.equal(_types(out["atom_types"]), _types(s["ref_atom_types"]))


def test_sample_full_size_1000_steps_merged_tiles(gold_full, monkeypatch):
    """the same 1000-step reference golden with the per-edge GEMMs FORCED onto the merged 128x256 single-accumulator
    tiles (the format the benchmark batch uses; 4 crystals would not select it): 2 000 chained evaluations of it"""
    from matinvent_b200.models.diffcsp import TapeNoise
    from oracle.ref_import import make_batch
    monkeypatch.setenv("MI_TC_FORCE_MERGED", "1")
    s = gold_full["sample_T1000"]
    m = _full_module(gold_full)
    assert m.decoder.edge_mode(int((s["num_atoms"] ** 2).sum())) == (True, True)
    out, _ = m.sample(make_batch(s["num_atoms"].tolist()), step_lr=s["step_lr"], noise=TapeNoise("cuda", seed=s["seed"]))
    print("1000-step parity (merged tiles forced): frac %.2e  lattice %.2e" % (
        wrapped_err(out["frac_coords"], s["ref_frac_coords"]), rel_err(out["lattices"], s["ref_lattices"])))
    assert wrapped_err(out["frac_coords"], s["ref_frac_coords"]) < 1e-4
    assert rel_err(out["lattices"], s["ref_lattices"]) < 1e-4
    assert torch.equal(_types(out["atom_types"]), _types(s["ref_atom_types"]))


def test_sample_full_size_1000_steps_ffma_path(gold_full):
    """Same case with the tensor-core GEMMs disabled (FP32 CUDA-core path) — kept as the accuracy yardstick."""
    from matinvent_b200.models.diffcsp import TapeNoise
    from oracle.ref_import import make_batch
    s = gold_full["sample_T1000"]
    m = _full_module(gold_full)
    m.decoder.use_tc = False
    out, _ = m.sample(make_batch(s["num_atoms"].tolist()), step_lr=s["step_lr"], noise=TapeNoise("cuda", seed=s["seed"]))
    print("1000-step parity (FFMA): frac %.2e  lattice %.2e" % (wrapped_err(out["frac_coords"], s["ref_frac_coords"]),
                                                                rel_err(out["lattices"], s["ref_lattices"])))
    assert wrapped_err(out["frac_coords"], s["ref_frac_coords"]) < 1e-4
    assert rel_err(out["lattices"], s["ref_lattices"]) < 1e-4
    assert torch.equal(_types(out["atom_types"]), _types(s["ref_atom_types"]))


def test_generate_plugin_api(gold_small):
    """DiffCSPSampler.generate (models/diffcsp/sample.py:148-201) through the plugin call, two batches on one noise tape:
    every returned crystal against the oracle's sampler + post-processing (argmax + 1, lattice -> lengths / angles,
    per-crystal split) on the same atom-count draw and the same tape"""
    import numpy as np
    from oracle import diffcsp_oracle as O
    from matinvent_b200.models.diffcsp import DiffCSPSampler, TapeNoise
    from matinvent_b200.models.diffcsp.sample import ATOM_DIST, DEFAULT_STEP_LR
    gs = gold_small
    hp, sd = gs["hp"], gs["sd"]
    m = build_module(hp, sd, gs["sigmas_norm"])
    np.random.seed(0)
    data, strucs = DiffCSPSampler(batch_size=6, num_batches=2).generate(m, noise=TapeNoise("cuda", seed=21), filter=None,
                                                                        max_num=3)
    assert len(data) == len(strucs) == 12
    np.random.seed(0)
    na = np.random.choice(len(ATOM_DIST["mp_20"]), 12, p=ATOM_DIST["mp_20"]).tolist()
    noise = O.Noise(torch.Generator().manual_seed(21))
    sch = O.Schedules(hp, gs["sigmas_norm"])
    ref = []
    for b in range(2):
        out = O.sample(sd, hp, sch, na[6 * b:6 * b + 6], noise, step_lr=DEFAULT_STEP_LR["gen"]["mp_20"])
        ref += O.generate_postprocess(out)
    for d, r in zip(data, ref):
        n = int(d.num_atoms)
        assert n == r["num_atoms"] and d.frac_coords.shape == (n, 3) and d.atom_types.shape == (n,)
        assert d.lengths.shape == (1, 3) and d.angles.shape == (1, 3)
        assert torch.equal(d.atom_types.long(), r["atom_types"].long())
        assert wrapped_err(d.frac_coords, r["frac_coords"]) < 1e-4
        assert rel_err(d.lengths, r["lengths"]) < 1e-4
        assert float((d.angles - r["angles"]).abs().max()) < 1e-2          # degrees
        assert 1 <= int(d.atom_types.min()) and int(d.atom_types.max()) <= 100


# ------------------------------------------------------------------------------------ knn edges
def _edge_multiset(src, dst, vec):
    key = torch.stack([src.double(), dst.double(), *(torch.round(vec.double() * 1e4).unbind(1))], dim=1)
    return sorted(map(tuple, key.tolist()))


def test_radius_graph_pbc_matches_reference(gold_rg):
    """Kernel output (symmetrised, grouped by source) vs the reference's radius_graph_pbc lists pushed through
    the oracle's reorder_symmetric_edges, compared as multisets of (src, dst, image offset)."""
    from oracle import diffcsp_oracle as O
    from matinvent_b200.models.diffcsp.knn import KnnGraph
    rg = gold_rg
    na = rg["num_atoms"]
    for K in (4, 20):
        ref = rg["ref_K%d" % K]
        ei, cell = ref["edge_index"].long(), ref["cell"].float()
        e_new, off_new, nb_new, _ = O.reorder_symmetric_edges(ei, cell, ref["per_image"], cell)
        g = KnnGraph(na.tolist(), "cuda", K)
        g.rebuild(rg["frac_coords"].cuda(), rg["lattices"].cuda())
        assert g.E == e_new.shape[1], (K, g.E, e_new.shape[1])
        E = g.E
        # reference frac_diff sign convention: gen_edges returns -vector, vector = x_j - x_i + offset for the
        # kept direction; in (src, dst, off) form the offset is -off_new (see mi_graph.cu phase 2)
        mine = _edge_multiset(g.edge_src[:E].cpu(), g.edge_dst[:E].cpu(), g.cell_off[:E].cpu())
        theirs = _edge_multiset(e_new[0], e_new[1], -off_new)
        assert mine == theirs, K
        sp = g.seg_ptr.cpu().long()
        assert int(sp[-1]) == E and torch.all(sp[1:] >= sp[:-1])
        for i in (0, 5, g.N - 1):
            assert torch.all(g.edge_src[sp[i]:sp[i + 1]].cpu() == i)
        dp, perm = g.dst_ptr.cpu().long(), g.dst_perm[:E].cpu().long()
        assert sorted(perm.tolist()) == list(range(E))
        for i in (0, 7, g.N - 1):
            assert torch.all(g.edge_dst[:E].cpu()[perm[dp[i]:dp[i + 1]]] == i)


def test_forward_small_knn(gold_small):
    gs = gold_small
    m = build_module(gs["hp_knn"], gs["sd"], gs["sigmas_norm"])
    c = gs["forward_knn"]
    pl, px, pt = _forward_case(m, c)
    g = m.decoder.graph_for(c["num_atoms"])
    E = g.E
    assert E == c["ref_edges"].shape[1]
    fd = torch.empty(E, 3, device="cuda")
    from matinvent_b200 import ops
    phi = torch.empty(E, 6 * gs["hp_knn"]["num_freqs"], device="cuda")
    ops.edge_fourier(c["x"].cuda(), g.edge_src, g.edge_dst, g.cell_off, E, gs["hp_knn"]["num_freqs"], fd, phi)
    mine = _edge_multiset(g.edge_src[:E].cpu(), g.edge_dst[:E].cpu(), fd.cpu())
    theirs = _edge_multiset(c["ref_edges"][0].long(), c["ref_edges"][1].long(), c["ref_frac_diff"])
    assert mine == theirs
    assert rel_err(pl, c["ref_pred_l"]) < 2e-5
    assert rel_err(px, c["ref_pred_x"]) < 2e-5
    assert rel_err(pt, c["ref_pred_t"]) < 2e-5


def test_knn_gradients_vs_oracle_autograd(gold_small):
    """No reference golden for knn gradients: compare with float32 autograd through the oracle restatement."""
    from oracle import diffcsp_oracle as O
    gs = gold_small
    hp = gs["hp_knn"]
    m = build_module(hp, gs["sd"], gs["sigmas_norm"])
    c = gs["forward_knn"]
    na = c["num_atoms"]
    n2g = torch.repeat_interleave(torch.arange(len(na)), na)
    pl, px, pt = m.decoder(c["temb"].cuda(), c["a"].cuda(), c["x"].cuda(), c["l"].cuda(), na, n2g)
    gen = torch.Generator().manual_seed(5)
    wl, wx, wt = (torch.randn(t.shape, generator=gen) for t in (pl, px, pt))
    ((pl * wl.cuda()).sum() + (px * wx.cuda()).sum() + (pt * wt.cuda()).sum()).backward()
    sd = {k: v.clone().double().requires_grad_(True) for k, v in gs["sd"].items()}
    ol, ox, ot = O.cspnet_forward(sd, hp, c["temb"].double(), c["a"].double(), c["x"].double(), c["l"].double(), na, n2g)
    ((ol * wl).sum() + (ox * wx).sum() + (ot * wt).sum()).backward()
    grads = m.decoder.reference_named_grads()
    worst = max((rel_err(grads[k], v.grad), k) for k, v in sd.items())
    assert worst[0] < 1e-4, worst


# ---------------------------------------------------------------- BASELINE batch (256 mp_20 crystals, ~34k edges)
# The golden cases above hold 4 crystals (601 edges): the per-edge GEMMs take the two-accumulator 128x128 tiles there.
# At the benchmark's batch they switch to the single-accumulator 128x256 tiles (CSPNet.edge_mode), so that path is
# pinned here: one forward against the oracle on the same inputs, gradients and a reverse trajectory against the
# FP32 CUDA-core path, which the golden cases above pin to the unmodified reference.
def _baseline_batch(seed=0):
    import numpy as np
    from matinvent_b200.models.diffcsp.sample import ATOM_DIST
    na = torch.tensor(np.random.RandomState(0).choice(21, 256, p=ATOM_DIST["mp_20"]).tolist())
    g = torch.Generator().manual_seed(seed)
    N, B = int(na.sum()), len(na)
    t = torch.randn(B, 256, generator=g)
    a = torch.randn(N, 100, generator=g)
    x = torch.rand(N, 3, generator=g)
    l = torch.randn(B, 3, 3, generator=g) + 4.0 * torch.eye(3)
    return na, t, a, x, l


def test_forward_baseline_batch_vs_oracle(gold_full):
    from oracle import diffcsp_oracle as O
    m = _full_module(gold_full)
    na, t, a, x, l = _baseline_batch()
    n2g = torch.repeat_interleave(torch.arange(len(na)), na)
    dec = m.decoder
    assert dec.edge_mode(int((na * na).sum())) == (True, True)          # merged tiles are what runs here
    with torch.no_grad():
        pl, px, pt = dec(t.cuda(), a.cuda(), x.cuda(), l.cuda(), na, n2g)
    sd = O.init_params(gold_full["hp"], gold_full["seeds"][0])
    with torch.no_grad():
        rl, rx, rt = O.cspnet_forward(sd, gold_full["hp"], t, a, x, l, na, n2g)
    errs = (rel_err(pl, rl), rel_err(px, rx), rel_err(pt, rt))
    print("baseline-batch forward vs oracle: lattice %.2e coord %.2e type %.2e" % errs)
    assert max(errs) < 2e-5, errs


def test_forward_baseline_batch_vs_reference_golden(gold_full, gold_baseline):
    """the same size against the UNMODIFIED reference: tests/golden/baseline_forward.pt (oracle/make_golden.py --baseline)"""
    from conftest import baseline_inputs
    m = _full_module(gold_full)
    na, t, a, x, l, n2g = baseline_inputs(gold_baseline)
    assert m.decoder.edge_mode(int((na * na).sum())) == (True, True)
    with torch.no_grad():
        pl, px, pt = m.decoder(t.cuda(), a.cuda(), x.cuda(), l.cuda(), na, n2g)
    gb = gold_baseline
    errs = (rel_err(pl, gb["ref_pred_l"]), rel_err(px, gb["ref_pred_x"]), rel_err(pt[::8], gb["ref_pred_t_rows8"]))
    print("baseline-batch forward vs reference golden: lattice %.2e coord %.2e type %.2e" % errs)
    assert max(errs) < 2e-5, errs


def test_ft_gradients_baseline_batch_vs_reference_golden(gold_full, gold_baseline):
    """one fine-tune timestep of the 256-crystal batch (reward-weighted loss + KL proxy) against the gradients of the
    UNMODIFIED reference's autograd: at this size the input- and weight-gradient GEMMs of the per-edge blocks run on
    the tensor cores (transposed operands, split-K accumulate)"""
    from matinvent_b200.models.diffcsp import TapeNoise
    from oracle.ref_import import make_batch
    ft = gold_baseline["ft"]
    agent, prior = _full_module(gold_full, 0), _full_module(gold_full, 1)
    for p in prior.parameters():
        p.requires_grad = False
    assert int((ft["num_atoms"] ** 2).sum()) >= agent.decoder.WGRAD_TC_ROWS
    batch = make_batch(ft["num_atoms"].tolist(), **ft["crystals"])
    batch.reward = batch.reward.cuda()
    noised = agent.add_noise(batch, ft["t_idx"], noise=TapeNoise("cuda", seed=ft["noise_seed"]))
    sample_loss, agent_pred = agent.calc_sample_loss(noised)
    _, prior_pred = prior.calc_sample_loss(noised)
    kl = agent.calc_kl_reg(agent_pred, prior_pred, batch)
    ((batch.reward * sample_loss + kl * (1.1 - batch.reward) * ft["sigma"]).mean() / ft["accum"]).backward()
    assert rel_err(sample_loss, ft["ref_sample_loss"]) < 1e-4 and rel_err(kl, ft["ref_kl"]) < 1e-4
    grads = agent.decoder.reference_named_grads()
    worst = 0.0
    for k, chk in gold_baseline["ft_grad_checks"].items():
        g = grads[k].cpu()
        assert abs(float(g.double().abs().sum()) - chk["abs_sum"]) < 1e-4 * chk["abs_sum"] + 1e-12, k
        scale = float(g.abs().max()) + 1e-30
        worst = max(worst, float((g.reshape(-1)[:64] - chk["head"]).abs().max()) / scale)
        assert float((g.reshape(-1)[:64] - chk["head"]).abs().max()) < 1e-4 * scale, k
    print("baseline-batch fine-tune gradients vs reference golden: worst head error %.2e of the tensor's max" % worst)
    worst = _check_grad_projections(grads, gold_baseline["ft_grad_proj"])
    print("baseline-batch fine-tune gradients vs reference: worst projection error %.2e of the tensor's norm (%s)" % worst)


def test_sample_baseline_batch_vs_reference_trajectory():
    """The benchmark configuration's multi-step behaviour against the UNMODIFIED reference: a complete 100-step
    DiffCSPModule.sample (the reference's own T=100 schedules) of the 256-crystal batch under a shared noise tape
    (tests/golden/baseline_traj.pt, oracle/make_golden.py --baseline-traj).  200 chained score-network evaluations on
    the merged 128x256 tiles (and the fused scatter epilogue), north-star tolerances."""
    from conftest import load_gold
    from oracle import diffcsp_oracle as O
    from oracle.ref_import import make_batch
    from matinvent_b200.models.diffcsp import TapeNoise
    gt = load_gold("baseline_traj.pt")
    hp = gt["hp"]
    sd = O.init_params(hp, gt["seed_weights"])
    for k, v in sd.items():
        assert abs(float(v.double().abs().sum()) - gt["checksums"][k]) <= 1e-9 * max(1.0, gt["checksums"][k]), k
    m = build_module(hp, sd, gt["sigmas_norm"])
    na = gt["num_atoms"]
    assert m.decoder.edge_mode(int((na * na).sum())) == (True, True)
    out, traj = m.sample(make_batch(na.tolist()), step_lr=gt["step_lr"], noise=TapeNoise("cuda", seed=gt["seed"]),
                         return_traj=True)
    for t, ref in gt["ref_traj"].items():
        ef, el = wrapped_err(traj[t]["frac_coords"], ref["frac_coords"]), rel_err(traj[t]["lattices"], ref["lattices"])
        print("baseline-batch trajectory vs reference, step %3d: frac %.2e lattice %.2e" % (t, ef, el))
        assert ef < 1e-4 and el < 1e-4, t
    ef, el = wrapped_err(out["frac_coords"], gt["ref_frac_coords"]), rel_err(out["lattices"], gt["ref_lattices"])
    print("baseline-batch 100-step sample vs reference: frac %.2e lattice %.2e" % (ef, el))
    assert ef < 1e-4 and el < 1e-4
    assert torch.equal(out["atom_types"].argmax(-1).cpu().to(torch.int8), gt["ref_types_argmax"])
    assert rel_err(out["atom_types"][::8], gt["ref_atom_types_rows8"]) < 1e-4


def test_gradients_baseline_batch_merged_vs_ffma(gold_full):
    """training forward (pre-activation stores) + hand-written backward at the benchmark's batch: tensor-core
    (merged tiles) against the FP32 CUDA-core path"""
    na, t, a, x, l = _baseline_batch(1)
    n2g = torch.repeat_interleave(torch.arange(len(na)), na)
    grads = []
    for use_tc in (True, False):
        m = _full_module(gold_full)
        m.decoder.use_tc = use_tc
        pl, px, pt = m.decoder(t.cuda(), a.cuda(), x.cuda(), l.cuda(), na, n2g)
        g = torch.Generator().manual_seed(5)
        wl, wx, wt = (torch.randn(p.shape, generator=g).cuda() for p in (pl, px, pt))
        ((pl * wl).sum() + (px * wx).sum() + (pt * wt).sum()).backward()
        grads.append({k: v.clone() for k, v in m.decoder.reference_named_grads().items()})
    worst = max((rel_err(grads[0][k], grads[1][k]), k) for k in grads[0])
    print("baseline-batch gradients, merged tensor-core vs FFMA: worst %.2e (%s)" % worst)
    assert worst[0] < 1e-4, worst


def test_sample_baseline_batch_merged_vs_ffma(gold_full):
    """60 reverse steps (end of the schedule, where the score network matters most) of the 256-crystal batch on a
    shared noise tape: tensor-core (merged tiles) against the FP32 CUDA-core path, same 1e-4 bar as the golden case"""
    from matinvent_b200.models.diffcsp import TapeNoise
    from oracle.ref_import import make_batch
    na = _baseline_batch()[0]
    outs = []
    for use_tc in (True, False):
        m = _full_module(gold_full)
        m.decoder.use_tc = use_tc
        out, _ = m.sample(make_batch(na.tolist()), step_lr=1e-5, noise=TapeNoise("cuda", seed=11), timesteps=60)
        outs.append(out)
    ef, el = wrapped_err(outs[0]["frac_coords"], outs[1]["frac_coords"]), rel_err(outs[0]["lattices"], outs[1]["lattices"])
    print("baseline-batch 60-step sample, merged tensor-core vs FFMA: frac %.2e lattice %.2e" % (ef, el))
    assert ef < 1e-4 and el < 1e-4
    assert torch.equal(_types(outs[0]["atom_types"]), _types(outs[1]["atom_types"]))


# ---------------------------------------------------------------- ragged / degenerate batches against the oracle
@pytest.mark.parametrize("num_atoms", [[1], [1, 1, 1], [20, 20, 20], [1, 20, 1, 7], [2] * 65, [20] * 33])
def test_forward_ragged_batches_vs_oracle(gold_small, num_atoms):
    """single-atom crystals (one self edge each), the maximum of the mp_20 prior (20 atoms), a one-crystal batch and
    batches whose node / edge counts straddle the 128-row tiles: forward against the oracle on the same inputs"""
    from oracle import diffcsp_oracle as O
    gs = gold_small
    hp, sd = gs["hp"], gs["sd"]
    m = build_module(hp, sd, gs["sigmas_norm"])
    na = torch.tensor(num_atoms)
    g = torch.Generator().manual_seed(17 + len(num_atoms))
    N, B = int(na.sum()), len(na)
    t = torch.randn(B, hp["time_dim"], generator=g)
    a = torch.randn(N, 100, generator=g)
    x = torch.rand(N, 3, generator=g)
    l = torch.randn(B, 3, 3, generator=g) + 4.0 * torch.eye(3)
    n2g = torch.repeat_interleave(torch.arange(B), na)
    with torch.no_grad():
        pl, px, pt = m.decoder(t.cuda(), a.cuda(), x.cuda(), l.cuda(), na, n2g)
        rl, rx, rt = O.cspnet_forward(sd, hp, t, a, x, l, na, n2g)
    assert pl.shape == rl.shape and px.shape == rx.shape and pt.shape == rt.shape
    assert max(rel_err(pl, rl), rel_err(px, rx), rel_err(pt, rt)) < 2e-5


@pytest.mark.parametrize("num_atoms", [[1], [1, 20, 1]])
def test_sample_degenerate_batches_vs_oracle(gold_small, num_atoms):
    """a 12-step reverse trajectory of degenerate batches on a shared noise tape, CUDA-graph path, against the oracle"""
    from matinvent_b200.models.diffcsp import TapeNoise
    from oracle import diffcsp_oracle as O
    from oracle.ref_import import make_batch
    gs = gold_small
    hp, sd = gs["hp"], gs["sd"]
    m = build_module(hp, sd, gs["sigmas_norm"])
    out, _ = m.sample(make_batch(num_atoms), step_lr=1e-5, noise=TapeNoise("cuda", seed=3), timesteps=12)
    ref = O.sample(sd, hp, O.Schedules(hp, gs["sigmas_norm"]), num_atoms, O.Noise(torch.Generator().manual_seed(3)),
                   step_lr=1e-5, timesteps=12)
    assert wrapped_err(out["frac_coords"], ref["frac_coords"]) < 1e-4
    assert rel_err(out["lattices"], ref["lattices"]) < 1e-4
    assert torch.equal(_types(out["atom_types"]), _types(ref["atom_types"]))


@pytest.mark.parametrize("style", ["knn", "fc"])
def test_forward_benchmark_kernels_small_batches_vs_oracle(style):
    """hidden 512 / 128 frequencies with the CTA-pair per-edge kernels and the cluster node chain forced at small edge
    counts (decoder.force_merged), fully-connected AND periodic k-nearest-neighbour graphs (arbitrary dst rows in the
    gathers, variable-length source segments in the fused scatter-mean), against the oracle"""
    from oracle import diffcsp_oracle as O
    hp = O.default_hparams(num_layers=2, edge_style=style, cutoff=6.0, max_neighbors=12, timesteps=10)
    sd = O.init_params(hp, seed=3)
    m = build_module(hp, sd, None)
    m.decoder.force_merged = True
    na = torch.tensor([5, 17, 1, 20, 9, 12, 3])
    B, N = len(na), int(na.sum())
    g = torch.Generator().manual_seed(4)
    t = torch.randn(B, 256, generator=g)
    a = torch.randn(N, 100, generator=g)
    x = torch.rand(N, 3, generator=g)
    l = 0.3 * torch.randn(B, 3, 3, generator=g) + 5.0 * torch.eye(3)
    n2g = torch.repeat_interleave(torch.arange(B), na)
    with torch.no_grad():
        out = m.decoder(t.cuda(), a.cuda(), x.cuda(), l.cuda(), na, n2g)
        ref = O.cspnet_forward(sd, hp, t, a, x, l, na, n2g)
    errs = [rel_err(u, r) for u, r in zip(out, ref)]
    print("benchmark kernels, %s graph, %d atoms: %s" % (style, N, " ".join("%.2e" % e for e in errs)))
    assert max(errs) < 2e-5, errs


def test_time_term_table_matches_per_step_linear(gold_full):
    """CSPNet.time_term_table (the per-crystal time term `temb W_t^T + b` of every timestep, one GEMM per sampling run) holds
    exactly the rows the per-step GEMM of forward_graph writes into ws.tb, and equals the float64 product to FP32-grade
    accuracy; mi_sampler_step_begin copies row t for every crystal (cspnet.py:267-271, diffusion.py:53-66)."""
    from matinvent_b200 import ops
    m = _full_module(gold_full)
    dec = m.decoder
    H = dec.hidden_dim
    ttab = m.time_table()                                     # [T + 1, time_dim]
    a = torch.randn(37, 100, device="cuda")
    table = dec.time_term_table(ttab, a)
    assert table.shape == (ttab.shape[0], H)
    bias = dec._emb_bias if dec._composed(a, False) else dec.w("lat_b")
    want = ttab.double() @ dec.w("lat_w_t").double().t() + bias.double()
    assert rel_err(table, want.float()) < 3e-6
    # the per-step path: B crystals at one timestep through the same linear
    B = 5
    for t in (0, 1, 499, 1000):
        temb = ttab[t].expand(B, -1).contiguous()
        tb = torch.empty(B, H, device="cuda")
        dec._linear(temb, "lat_w_t", tb, B, bias=bias)
        assert torch.equal(tb, table[t].expand(B, -1)), t
        out = torch.full((B, H), 7.0, device="cuda")
        ops.sampler_step_begin(torch.tensor([t], dtype=torch.int32, device="cuda"), table, out, B, H)
        assert torch.equal(out, table[t].expand(B, -1))
