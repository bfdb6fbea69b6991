"""Generates tests/golden/*.pt FROM THE UNMODIFIED REFERENCE (imported under oracle/shims; see
oracle/ref_import.py).  Run in the build container only:  python -m oracle.make_golden [--full]

Every tensor stored under a key starting with "ref_" was produced by /root/reference code
(models/diffcsp/{cspnet,diffusion,scheduler,utils}.py); the oracle restatement and the CUDA path are
both tested against them.  Noise is drawn from CPU torch.Generator tapes in the reference's own
draw order by temporarily replacing torch.rand / randn / randn_like.
Weights: oracle.diffcsp_oracle.init_params(hp, seed) (default nn.Linear init, heads x0.05); small
nets are stored, the full-size net is regenerated from its seed and pinned by per-tensor checksums.
"""
import argparse
import contextlib
import os
import sys
import time

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import diffcsp_oracle as O  # noqa: E402
from oracle import ref_import as R  # noqa: E402

GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")


@contextlib.contextmanager
def noise_tape(seed):
    gen = torch.Generator().manual_seed(seed)
    orig = (torch.rand, torch.randn, torch.randn_like)
    torch.rand = lambda s, **k: orig[0](tuple(s), generator=gen)
    torch.randn = lambda s, **k: orig[1](tuple(s), generator=gen)
    torch.randn_like = lambda x: orig[1](tuple(x.shape), generator=gen, dtype=x.dtype)
    try:
        yield gen
    finally:
        torch.rand, torch.randn, torch.randn_like = orig


def checksums(sd):
    return {k: float(v.double().abs().sum()) for k, v in sd.items()}


def synth_crystals(num_atoms, seed):
    g = torch.Generator().manual_seed(seed)
    B, N = len(num_atoms), int(sum(num_atoms))
    return dict(lengths=3 + 5 * torch.rand(B, 3, generator=g), angles=70 + 40 * torch.rand(B, 3, generator=g),
                frac_coords=torch.rand(N, 3, generator=g), atom_types=torch.randint(1, 101, (N,), generator=g),
                reward=torch.rand(B, generator=g))


def forward_case(hp, sd, ref, num_atoms, seed, t_int):
    na = torch.tensor(num_atoms)
    B, N = len(num_atoms), int(na.sum())
    n2g = torch.repeat_interleave(torch.arange(B), na)
    g = torch.Generator().manual_seed(seed)
    t = O.time_embedding(torch.full((B,), t_int), hp["time_dim"])
    a = torch.randn(N, 100, generator=g)
    x = torch.rand(N, 3, generator=g)
    l = torch.randn(B, 3, 3, generator=g)
    if hp["edge_style"] == "knn":
        l = torch.eye(3)[None] * 5 + 0.5 * l
    with torch.no_grad():
        pl, px, pt = ref.decoder(t, a, x, l, na, n2g)
        e, fd = ref.decoder.gen_edges(na, x, l, n2g)
    return dict(num_atoms=na, t_int=t_int, temb=t, a=a, x=x, l=l, ref_pred_l=pl, ref_pred_x=px, ref_pred_t=pt,
                ref_edges=e.to(torch.int32), ref_frac_diff=fd)


def build(hp, seed_agent=0, seed_prior=1, sigmas_norm=None):
    torch.manual_seed(1234)   # SigmaScheduler's Monte-Carlo sigma_norm draw (scheduler.py:46-51)
    ref = R.build_reference_module(hp, sigmas_norm)
    prior = R.build_reference_module(hp, ref.sigma_scheduler.sigmas_norm)
    sd, sdp = O.init_params(hp, seed_agent), O.init_params(hp, seed_prior)
    ref.decoder.load_state_dict(sd)
    prior.decoder.load_state_dict(sdp)
    return ref, prior, sd, sdp


def ft_case(hp, ref, prior, num_atoms, t_idx, sigma=0.025, accum=50, data_seed=3, noise_seed=11):
    cr = synth_crystals(num_atoms, data_seed)
    batch = R.make_batch(num_atoms, **cr)
    ref.zero_grad()
    with noise_tape(noise_seed):
        noised = ref.add_noise(batch, t_idx)
        sl, ap = ref.calc_sample_loss(noised)
        with torch.no_grad():
            _, pp = prior.calc_sample_loss(noised)
        kl = ref.calc_kl_reg(ap, pp, batch)
        loss = (batch.reward * sl + kl * (1.1 - batch.reward) * sigma).mean() / accum
        loss.backward()
    grads = {k: p.grad.detach().clone() for k, p in ref.decoder.named_parameters()}
    (temb, a_t, x_t, l_t, _, _), (rand_l, tar_x, rand_t), _ = noised
    return dict(crystals=cr, num_atoms=torch.tensor(num_atoms), t_idx=t_idx, sigma=sigma, accum=accum,
                noise_seed=noise_seed, ref_temb=temb, ref_a_t=a_t.detach(), ref_x_t=x_t.detach(), ref_l_t=l_t.detach(),
                ref_rand_l=rand_l, ref_tar_x=tar_x, ref_rand_t=rand_t, ref_sample_loss=sl.detach(),
                ref_kl=kl.detach(), ref_loss=loss.detach(), ref_agent_pred=[p.detach() for p in ap],
                ref_prior_pred=[p.detach() for p in pp]), grads


def sample_case(hp, ref, num_atoms, seed, step_lr=5e-6):
    batch = R.make_batch(num_atoms)
    t0 = time.time()
    with noise_tape(seed):
        out, traj = ref.sample(batch, step_lr=step_lr)
    T = hp["timesteps"]
    keep = sorted(set([T - 1, T - 2, T // 2, 1, 0]))
    return dict(num_atoms=torch.tensor(num_atoms), seed=seed, step_lr=step_lr, seconds=time.time() - t0,
                ref_frac_coords=out["frac_coords"], ref_lattices=out["lattices"], ref_atom_types=out["atom_types"],
                ref_traj={t: {k: traj[t][k] for k in ("frac_coords", "lattices")} for t in keep if t in traj})


def baseline_forward():
    """One forward of the UNMODIFIED reference CSPNet (full size) on the benchmark's batch: 256 crystals with the
    mp_20 atom-count prior drawn like bench.py does (34 445 edges) — the size at which the CUDA path switches its
    per-edge GEMMs to 128x256 single-accumulator tiles.  Inputs are regenerated from the stored seed by the tests;
    of the [N,100] type head every 8th row is kept (fixture size)."""
    import numpy as np
    hp = O.default_hparams()
    sn = torch.load(os.path.join(GOLD, "sigmas_norm_T1000.pt"))["sigmas_norm"]
    ref, prior, sd, sdp = build(hp, sigmas_norm=sn)
    # models/diffcsp/sample.py ATOM_DIST['mp_20'] of the reference (number-of-atoms prior), bench.py's draw
    # (read as a literal from the source: importing that module pulls in pymatgen)
    import ast
    src = open(os.path.join(R.REF_ROOT, "models", "diffcsp", "sample.py")).read()
    node = next(n for n in ast.parse(src).body if isinstance(n, ast.Assign) and
                getattr(n.targets[0], "id", "") == "ATOM_DIST")
    dist = ast.literal_eval(node.value)["mp_20"]
    num_atoms = np.random.RandomState(0).choice(len(dist), 256, p=dist).tolist()
    t0 = time.time()
    c = forward_case(hp, sd, ref, num_atoms, 101, 650)
    gold = dict(hp=hp, seed_weights=0, checksums=checksums(sd), num_atoms=c["num_atoms"], seed=101, t_int=650,
                ref_pred_l=c["ref_pred_l"], ref_pred_x=c["ref_pred_x"], ref_pred_t_rows8=c["ref_pred_t"][::8].clone(),
                edges=int((c["num_atoms"] ** 2).sum()), seconds=time.time() - t0)
    # one fine-tune timestep of the same batch (reward-weighted loss + KL proxy, all gradients by the reference's
    # autograd): the long reductions of the weight gradients are what the CUDA path moves to the tensor cores
    t1 = time.time()
    ft, grads = ft_case(hp, ref, prior, [max(1, n) for n in num_atoms], 300)
    keep = ("crystals", "num_atoms", "t_idx", "sigma", "accum", "noise_seed", "ref_sample_loss", "ref_kl", "ref_loss")
    gold["ft"] = {k: ft[k] for k in keep}
    gold["ft_grad_checks"] = {k: dict(abs_sum=float(v.double().abs().sum()), head=v.reshape(-1)[:64].clone())
                              for k, v in grads.items()}
    gold["ft_grad_proj"] = grad_projections(grads)
    gold["checksums_prior"] = checksums(sdp)
    gold["ft_seconds"] = time.time() - t1
    torch.save(gold, os.path.join(GOLD, "baseline_forward.pt"))
    print("baseline_forward.pt written: %d atoms, %d edges, reference forward %.1f s" %
          (int(c["num_atoms"].sum()), gold["edges"], gold["seconds"]))


def _bench_atom_counts(n):
    """bench.py's draw: np.random.RandomState(0).choice over the reference's ATOM_DIST['mp_20'] literal (read from the
    source: importing models/diffcsp/sample.py pulls in pymatgen)"""
    import ast
    import numpy as np
    src = open(os.path.join(R.REF_ROOT, "models", "diffcsp", "sample.py")).read()
    node = next(n_ for n_ in ast.parse(src).body if isinstance(n_, ast.Assign) and
                getattr(n_.targets[0], "id", "") == "ATOM_DIST")
    dist = ast.literal_eval(node.value)["mp_20"]
    return np.random.RandomState(0).choice(len(dist), n, p=dist).tolist()


def grad_projections(grads, n_proj=4):
    """position- and sign-sensitive fingerprints of a gradient set: for every tensor, its L2 norm and the inner products
    with `n_proj` seeded standard-normal tensors (seed = crc32 of the parameter name + projection index).  A transposed,
    permuted or sign-flipped block moves a projection by ~ the block's norm; |.|-sums and a 64-element head do not see it."""
    import zlib
    out = {}
    for k, v in grads.items():
        v64 = v.detach().double().reshape(-1)
        proj = []
        for j in range(n_proj):
            g = torch.Generator().manual_seed(zlib.crc32(("%s#%d" % (k, j)).encode()))
            r = torch.randn(v64.numel(), generator=g, dtype=torch.float64)
            proj.append(float(torch.dot(v64, r)))
        out[k] = dict(norm=float(v64.norm()), proj=proj)
    return out


def add_grad_projections():
    """re-run the two full-size fine-tune timesteps of the UNMODIFIED reference (4 crystals: full_net.pt, 256 crystals:
    baseline_forward.pt) and add `ft_grad_proj` to the existing fixtures (everything else in them is left as it is)"""
    hp = O.default_hparams()
    sn = torch.load(os.path.join(GOLD, "sigmas_norm_T1000.pt"))["sigmas_norm"]
    ref, prior, sd, sdp = build(hp, sigmas_norm=sn)
    for fname, num_atoms in (("full_net.pt", [4, 11, 20, 8]), ("baseline_forward.pt", [max(1, n) for n in _bench_atom_counts(256)])):
        gold = torch.load(os.path.join(GOLD, fname), weights_only=False)
        ft, grads = ft_case(hp, ref, prior, num_atoms, 300)
        assert torch.equal(ft["num_atoms"], gold["ft"]["num_atoms"])
        worst = 0.0
        for k, chk in gold["ft_grad_checks"].items():      # the re-run reproduces the stored fingerprints
            worst = max(worst, abs(float(grads[k].double().abs().sum()) - chk["abs_sum"]) / (chk["abs_sum"] + 1e-300))
        assert worst < 1e-6, worst
        gold["ft_grad_proj"] = grad_projections(grads)
        torch.save(gold, os.path.join(GOLD, fname))
        print("%s: ft_grad_proj added (%d tensors; re-run vs stored abs-sums: %.1e)" % (fname, len(grads), worst))


def baseline_trajectory(T=100, seed=7):
    """A complete T-step `DiffCSPModule.sample` of the UNMODIFIED reference (full-size net) on the benchmark's batch of
    256 mp_20 crystals (34 445 edges) under a noise tape: the multi-step behaviour of the configuration bench.py
    measures, where the CUDA path runs its per-edge GEMMs on the merged 128x256 single-accumulator tiles.  The
    schedules are the reference's own for `timesteps=T` (the whole sigma / beta range in T steps); its Monte-Carlo
    sigmas_norm buffer is stored.  Kept: the final state (atom types as argmax int8 + every 8th row of the
    continuous state), coordinates and lattices at four intermediate steps."""
    hp = O.default_hparams(timesteps=T)
    ref, prior, sd, sdp = build(hp)
    num_atoms = _bench_atom_counts(256)
    batch = R.make_batch(num_atoms)
    t0 = time.time()
    with noise_tape(seed):
        out, traj = ref.sample(batch, step_lr=5e-6)
    keep = sorted(set([T - 1, (3 * T) // 4, T // 2, T // 4, 1]))
    gold = dict(hp=hp, seed_weights=0, checksums=checksums(sd), num_atoms=torch.tensor(num_atoms), seed=seed, step_lr=5e-6,
                sigmas_norm=ref.sigma_scheduler.sigmas_norm.clone(), seconds=time.time() - t0,
                ref_frac_coords=out["frac_coords"].clone(), ref_lattices=out["lattices"].clone(),
                ref_types_argmax=out["atom_types"].argmax(-1).to(torch.int8),
                ref_atom_types_rows8=out["atom_types"][::8].clone(),
                ref_traj={t: {k: traj[t][k].clone() for k in ("frac_coords", "lattices")} for t in keep})
    torch.save(gold, os.path.join(GOLD, "baseline_traj.pt"))
    print("baseline_traj.pt written: %d crystals, %d steps, reference sample %.1f s" % (len(num_atoms), T, gold["seconds"]))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--full", action="store_true", help="also the full-size net incl. a 1000-step sample (minutes)")
    ap.add_argument("--baseline", action="store_true",
                    help="only: one forward of the full-size net on the benchmark batch (256 mp_20 crystals)")
    ap.add_argument("--baseline-traj", action="store_true",
                    help="only: a complete 100-step sample of the full-size net on the benchmark batch (minutes)")
    ap.add_argument("--grad-projections", action="store_true",
                    help="only: add seeded random projections of the full-size reference gradients to the fixtures")
    args = ap.parse_args()
    os.makedirs(GOLD, exist_ok=True)
    if args.baseline:
        return baseline_forward()
    if args.baseline_traj:
        return baseline_trajectory()
    if args.grad_projections:
        return add_grad_projections()

    # --- Monte-Carlo sigmas_norm buffer for T=1000 (copied, never recomputed: SURVEY.md §7 hard parts)
    hp_full = O.default_hparams()
    torch.manual_seed(1234)
    _, _, S, U = R.import_reference()
    sn1000 = S.SigmaScheduler(1000, hp_full["sigma_begin"], hp_full["sigma_end"]).sigmas_norm.clone()
    torch.save(dict(sigmas_norm=sn1000, sigma_begin=hp_full["sigma_begin"], sigma_end=hp_full["sigma_end"]),
               os.path.join(GOLD, "sigmas_norm_T1000.pt"))

    # --- small net, fc + knn: forward, ft timestep with all grads, T=40 sample
    hp = O.default_hparams(hidden_dim=128, num_layers=2, num_freqs=16, timesteps=40)
    ref, prior, sd, sdp = build(hp)
    num_atoms = [3, 1, 7, 20, 5, 12]
    gold = dict(hp=hp, sd=sd, sd_prior=sdp, sigmas_norm=ref.sigma_scheduler.sigmas_norm.clone(),
                beta=dict(ref.beta_scheduler.named_buffers()), sigma_sigmas=ref.sigma_scheduler.sigmas.clone())
    gold["forward_fc"] = forward_case(hp, sd, ref, num_atoms, 1, 17)
    hp_knn = dict(hp, edge_style="knn", max_neighbors=6)
    ref.decoder.edge_style, ref.decoder.max_neighbors = "knn", 6
    gold["hp_knn"] = hp_knn
    gold["forward_knn"] = forward_case(hp_knn, sd, ref, num_atoms, 2, 9)
    ref.decoder.edge_style, ref.decoder.max_neighbors = "fc", hp["max_neighbors"]
    gold["ft"], gold["ft_grads"] = ft_case(hp, ref, prior, num_atoms, 4)
    gold["sample"] = sample_case(hp, ref, num_atoms, 7)
    torch.save(gold, os.path.join(GOLD, "small_net.pt"))
    print("small_net.pt written")

    # --- radius_graph_pbc known answers (reference utils.py:335-601), K = 4 and 20
    g = torch.Generator().manual_seed(21)
    na = torch.tensor([2, 9, 20, 1, 14])
    cr = synth_crystals(na.tolist(), 22)
    lat = U.lattice_params_to_matrix_torch(cr["lengths"], cr["angles"])
    cart = torch.einsum('bi,bij->bj', cr["frac_coords"], lat[torch.repeat_interleave(torch.arange(5), na)])
    rg = dict(num_atoms=na, lattices=lat, frac_coords=cr["frac_coords"], cart=cart)
    for K in (4, 20):
        ei, uc, nb = U.radius_graph_pbc(cart, None, None, na, 7.0, K, device="cpu", lattices=lat)
        rg["ref_K%d" % K] = dict(edge_index=ei.to(torch.int32), cell=uc.to(torch.int8), per_image=nb)
    torch.save(rg, os.path.join(GOLD, "radius_graph.pt"))
    print("radius_graph.pt written")

    if args.full:
        hp = hp_full
        ref, prior, sd, sdp = build(hp, sigmas_norm=sn1000)
        num_atoms = [4, 11, 20, 8]
        gold = dict(hp=hp, seeds=(0, 1), checksums=checksums(sd), checksums_prior=checksums(sdp))
        gold["forward_fc"] = forward_case(hp, sd, ref, num_atoms, 1, 500)
        ft, grads = ft_case(hp, ref, prior, num_atoms, 300)
        gold["ft"] = ft
        gold["ft_grad_checks"] = {k: dict(abs_sum=float(v.double().abs().sum()), head=v.reshape(-1)[:64].clone())
                                  for k, v in grads.items()}
        gold["ft_grad_proj"] = grad_projections(grads)
        gold["sample_T1000"] = sample_case(hp, ref, num_atoms, 7)
        print("1000-step reference sample took %.1f s" % gold["sample_T1000"]["seconds"])
        torch.save(gold, os.path.join(GOLD, "full_net.pt"))
        print("full_net.pt written")


if __name__ == "__main__":
    main()
