"""GPU: the fused fine-tune step, the device replay buffer and the plugin/pipeline plumbing."""
import copy
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT, build_module, rel_err

pytestmark = pytest.mark.gpu


def _ft_batch(num_atoms, seed):
    from matinvent_b200.models.diffcsp.sample import CrystalBatch, CrystalData
    g = torch.Generator().manual_seed(seed)
    data = []
    for n in num_atoms:
        d = CrystalData(torch.rand(n, 3, generator=g), torch.randint(1, 101, (n,), generator=g),
                        3 + 5 * torch.rand(1, 3, generator=g), 70 + 40 * torch.rand(1, 3, generator=g), torch.tensor(n))
        d.reward = torch.rand(1, generator=g)
        data.append(d)
    return data, CrystalBatch(data)


def test_fused_ft_step_matches_oracle(gold_small):
    """2 epochs x 40 timesteps, Adam every 10 (pipeline/mat_invent.py:125-189) vs the oracle restatement with
    torch autograd + torch.optim.Adam on the CPU, shared noise tape."""
    from oracle import diffcsp_oracle as O
    from oracle.ref_import import make_batch
    from matinvent_b200.models.diffcsp import TapeNoise
    from matinvent_b200.models.diffcsp.finetune import FineTuner
    gs = gold_small
    hp = gs["hp"]
    num_atoms = [3, 9, 1, 14, 6]
    data, batch = _ft_batch(num_atoms, 4)
    agent = build_module(hp, gs["sd"], gs["sigmas_norm"])
    prior = build_module(hp, gs["sd_prior"], gs["sigmas_norm"])
    lr, accum, sigma, epochs, T = 1e-4, 10, 0.025, 2, hp["timesteps"]
    tuner = FineTuner(agent, prior, lr=lr, accum_steps=accum, sigma=sigma, noise=TapeNoise("cuda", seed=77))
    logs = [tuner.run_batch(batch, T) for _ in range(epochs)]
    # oracle
    ob = make_batch(num_atoms, lengths=batch.lengths, angles=batch.angles, frac_coords=batch.frac_coords,
                    atom_types=batch.atom_types)
    sd0 = {k: v.clone() for k, v in gs["sd"].items()}
    sch = O.Schedules(hp, gs["sigmas_norm"])
    noise = O.Noise(torch.Generator().manual_seed(77))
    sd1, ologs = O.ft_step({k: v.clone() for k, v in sd0.items()}, gs["sd_prior"], hp, sch, ob, batch.reward, noise,
                           lr=lr, accum_steps=accum, epochs=epochs, sigma=sigma, timesteps=T)
    for (l, d, k), (ol, od, ok) in zip(logs, ologs):
        B = len(num_atoms)
        assert abs(l - ol) < 1e-4 * abs(ol) and abs(d / B - od) < 1e-4 * abs(od) and abs(k / B - ok) < 1e-4 * abs(ok)
    mine = agent.decoder.state_dict()
    moved = diff_big = total = 0
    sum_abs = 0.0
    for k, v in sd1.items():
        dm, do = mine[k].cpu() - sd0[k], v.detach() - sd0[k]
        e = (dm - do).abs()
        total += e.numel()
        diff_big += int((e > 2e-6).sum())
        sum_abs += float(e.sum())
        moved += int((do.abs() > 1e-5).sum())
    assert moved > 0.5 * total              # Adam really moved the weights (8 steps of lr 1e-4)
    assert diff_big < 1e-3 * total, (diff_big, total)
    assert sum_abs / total < 1e-7


def test_fused_ft_equals_plugin_autograd_path(gold_small):
    """The fused engine and the autograd plugin path (what the unmodified reference loop would drive) give the
    same gradients for one timestep."""
    from matinvent_b200.models.diffcsp import TapeNoise
    from matinvent_b200.models.diffcsp.finetune import FineTuner
    gs = gold_small
    hp = gs["hp"]
    data, batch = _ft_batch([4, 11, 2], 9)
    a1 = build_module(hp, gs["sd"], gs["sigmas_norm"])
    a2 = build_module(hp, gs["sd"], gs["sigmas_norm"])
    prior = build_module(hp, gs["sd_prior"], gs["sigmas_norm"])
    for p in prior.parameters():
        p.requires_grad = False
    tuner = FineTuner(a1, prior, lr=1e-4, accum_steps=50, sigma=0.025, noise=TapeNoise("cuda", seed=5))
    tuner.optimizer_step = lambda: None      # keep the accumulated gradient
    tuner.run_batch(batch, 1)
    g1 = a1.decoder.flat_grad().clone()
    batch.to("cuda")
    noised = a2.add_noise(batch, 0, noise=TapeNoise("cuda", seed=5))
    sl, ap = a2.calc_sample_loss(noised)
    _, pp = prior.calc_sample_loss(noised)
    kl = a2.calc_kl_reg(ap, pp, batch)
    ((batch.reward * sl + kl * (1.1 - batch.reward) * 0.025).mean() / 50).backward()
    g2 = a2.decoder.flat_grad()
    assert rel_err(g1, g2) < 1e-5


def test_replay_buffer_matches_oracle():
    from oracle.diffcsp_oracle import ReplayBufferOracle, reduced_composition_key
    from matinvent_b200.memory import ReplayBuffer
    from matinvent_b200.models.diffcsp.sample import CrystalData
    rng = np.random.RandomState(0)
    buf = ReplayBuffer(buffer_size=12, sample_size=5, reward_cutoff=0.1, device="cuda")
    orc = ReplayBufferOracle(buffer_size=12, sample_size=5, reward_cutoff=0.1)
    serial = 0

    def make(n_items):
        nonlocal serial
        items = []
        for _ in range(n_items):
            n = int(rng.randint(1, 9))
            z = rng.choice([8, 8, 8, 14, 26, 3], size=n)           # few elements -> many duplicate compositions
            d = CrystalData(torch.rand(n, 3), torch.as_tensor(z), torch.rand(1, 3) + 3, torch.rand(1, 3) + 80, torch.tensor(n))
            d.serial = serial
            serial += 1
            items.append(d)
        return items

    for it in range(6):
        items = make(15)
        rewards = np.round(rng.rand(15), 1)                        # ties on purpose
        keys = [reduced_composition_key(d.atom_types.tolist()) for d in items]
        buf.extend(items, items, rewards)
        orc.extend(items, keys, rewards)
        assert len(buf) == len(orc)
        mine = sorted(np.round(buf.rewards.cpu().numpy().astype(float), 6).tolist(), reverse=True)
        theirs = sorted([round(float(np.float32(r[2])), 6) for r in orc.rows], reverse=True)
        assert mine == theirs, it
        # same compositions survive
        mk = sorted(reduced_composition_key(buf._rows["Z"][i, :int(buf._rows["n"][i])].tolist()) for i in range(len(buf)))
        assert mk == sorted(r[1] for r in orc.rows)
        if it == 3:
            purge = items[:4]
            buf.memory_purge(purge)
            orc.memory_purge([reduced_composition_key(d.atom_types.tolist()) for d in purge])
            assert len(buf) == len(orc)
    np.random.seed(3)
    data, rew = buf.sample()
    assert len(data) == min(len(buf), 5) and len(set(id(d) for d in data)) == len(data)
    for d, r in zip(data, rew):
        assert d.frac_coords.shape == (int(d.num_atoms), 3) and 0.1 < r <= 1.0
    assert ReplayBuffer(device="cuda").sample() == ([], [])
    # composition key = reduced formula class: Fe2O3 == Fe4O6 != Fe3O4
    k = buf.keys_of([torch.tensor([26, 26, 8, 8, 8]), torch.tensor([8] * 6 + [26] * 4), torch.tensor([26] * 3 + [8] * 4)])
    assert int(k[0]) == int(k[1]) and int(k[0]) != int(k[2])


def _suite(tmp_path, **over):
    from matinvent_b200.models.suite import DiffCSPSuite
    model = dict(decoder=dict(hidden_dim=64, num_layers=2, num_freqs=8), beta_scheduler=dict(timesteps=20, scheduler_mode="cosine"),
                 sigma_scheduler=dict(timesteps=20, sigma_begin=0.005, sigma_end=0.5))
    kw = dict(model_name="diffcsp", sample_cfg=dict(batch_size=8, num_batches=1),
              finetune_cfg=dict(batch_size=4, timesteps=20, lr=1e-4), device="cuda", random_init=True, model=model,
              head_scale=0.05)
    kw.update(over)
    return DiffCSPSuite(**kw)


def test_suite_checkpoint_roundtrip(tmp_path):
    from matinvent_b200.models.suite import DiffCSPSuite
    s = _suite(tmp_path)
    m = s.load_model()
    s.save_model(m, str(tmp_path / "ck"))
    assert os.path.isfile(tmp_path / "ck" / "last.ckpt") and os.path.isfile(tmp_path / "ck" / "hparams.yaml")
    blob = torch.load(tmp_path / "ck" / "last.ckpt", weights_only=False)
    assert "decoder.csp_layer_1.node_mlp.2.weight" in blob["state_dict"] and "config" in blob
    s2 = DiffCSPSuite(model_name="diffcsp", sample_cfg=dict(batch_size=8, num_batches=1),
                      finetune_cfg=dict(batch_size=4, timesteps=20, lr=1e-4), model_path=str(tmp_path / "ck"), device="cuda")
    m2 = s2.load_model()
    assert torch.equal(m.decoder.flat.data, m2.decoder.flat.data)
    assert torch.equal(m.sigma_scheduler.sigmas_norm, m2.sigma_scheduler.sigmas_norm)
    with pytest.raises(RuntimeError):
        DiffCSPSuite(model_name="diffcsp", sample_cfg={}, finetune_cfg={}, device="cuda").load_model()


def test_rl_loop_plumbing(tmp_path):
    """BASELINE configs[0]-style plumbing: small batch, short reverse process, stand-in composition reward,
    replay buffer, 2 RL iterations through MatInvent.run_rl; the agent moves, the prior does not."""
    from matinvent_b200.pipeline import MatInvent
    from matinvent_b200.pipeline.standin_reward import StandInHHIReward
    np.random.seed(0)
    torch.manual_seed(0)
    suite = _suite(tmp_path)
    # a random-init net blows the cells up (|l| ~ 1e2..1e5) and puts atoms anywhere: the device pre-filter runs with
    # thresholds that let these crystals through (the reference thresholds are tested in test_gpu_postsample.py)
    import functools
    from matinvent_b200.pipeline.filters import invalid_filter
    calls = []

    def loose_filter(data, strucs):
        calls.append(len(data))
        return invalid_filter(data, strucs, device="cuda", structure_validity=False, max_len=1e30)

    pipe = MatInvent(rl_epoch=2, model_suite=suite, reward=StandInHHIReward(),
                     sample_cfg=dict(filter=None, max_num=4, invalid_filter=loose_filter),
                     finetune_cfg=dict(batch_size=4, accum_steps=5, epochs=1, sigma=0.025), save_dir=str(tmp_path),
                     save_freq=1, device="cuda", replay=True, replay_args=dict(buffer_size=10, sample_size=2, reward_cutoff=0.0))
    w0 = pipe.agent.decoder.flat.data.clone()
    p0 = pipe.prior.decoder.flat.data.clone()
    pipe.run_rl()
    assert not torch.equal(pipe.agent.decoder.flat.data, w0) and torch.equal(pipe.prior.decoder.flat.data, p0)
    assert torch.isfinite(pipe.agent.decoder.flat.data).all()
    # max_num caps what is scored (pipeline/mat_invent.py:110-114): 2 iterations x min(8 sampled, 4) crystals
    assert len(pipe.replay) > 0 and pipe.cost == 8 and calls == [8, 8]
    assert os.path.isfile(tmp_path / "models" / "final" / "last.ckpt")
    # the reference's per-iteration dumps: valid / eval extxyz and the long-term-memory csv (:82-86, 117-121, 212)
    from matinvent_b200.pipeline.utils import read_extxyz
    assert len(read_extxyz(str(tmp_path / "samples" / "step_0001_valid.extxyz"))) == 8
    assert len(read_extxyz(str(tmp_path / "samples" / "step_0001_eval.extxyz"))) == 4
    assert os.path.isfile(tmp_path / "samples" / "long_term_memory.csv") and len(pipe.ltm) == 8
    assert set(pipe.timing) >= {"sample_s", "filter_s", "reward_s", "memory_s", "finetune_s", "total_s"}


def test_rl_step_device_reward_and_filter(tmp_path):
    """one RL iteration with every post-sampling stage on the device: validity pre-filter, multi-objective composition
    reward (min of two scaled properties, BASELINE configs[4] style), diversity filter, replay buffer"""
    from matinvent_b200.pipeline import MatInvent
    from matinvent_b200.pipeline.filters import invalid_filter
    from matinvent_b200.rewards import CompositionReward, synthetic_table
    np.random.seed(1)
    torch.manual_seed(1)
    suite = _suite(tmp_path, sample_cfg=dict(batch_size=32, num_batches=1))
    reward = CompositionReward(
        prop_cfg=[dict(name="hhi", table=synthetic_table("hhi"), target="descending", minv=750, maxv=3250),
                  dict(name="magmom", table=synthetic_table("magmom"), weights="atom", target="ascending", minv=0.0, maxv=0.25)],
        reward_threshold=0.8, reduce="min", device="cuda")
    pipe = MatInvent(rl_epoch=1, model_suite=suite, reward=reward,
                     sample_cfg=dict(invalid_filter=lambda d, s: invalid_filter(d, s, device="cuda", structure_validity=False,
                                                                                max_len=1e30)),
                     finetune_cfg=dict(batch_size=8, accum_steps=5, epochs=1, sigma=0.025), save_dir=str(tmp_path), save_freq=1,
                     device="cuda", replay=True, replay_args=dict(buffer_size=10, sample_size=2, reward_cutoff=0.0),
                     div_filter=True, df_args=dict(tol=3, buff=6))
    log, logs = pipe.rl_step()
    assert pipe.cost <= 32 and "hhi mean" in log and "magmom std" in log and log["crystal_num"] == len(pipe.ltm)
    assert len(logs) == 1 and np.isfinite(logs[0]["loss"])


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_ft_and_sampling_match_single_gpu(tmp_path):
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29611",
                          os.path.join(ROOT, "tests", "dist_check.py"), "--backend", "nccl"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "DIST_CHECK_OK" in out.stdout


def test_long_term_memory_keys_and_filter_on_device():
    """memory/ltm.py:30-109 through the public API: composition / element-set keys from the device kernel partition the
    crystals like pymatgen's reduced formula / element tuple would; the occurrence penalty follows the reference rule"""
    from matinvent_b200.memory import LongTimeMem
    from matinvent_b200.models.diffcsp.sample import CrystalData
    from oracle.ltm_oracle import LongTimeMemOracle
    import math
    from collections import Counter

    def crystal(z):
        n = len(z)
        return CrystalData(torch.rand(n, 3), torch.tensor(z), torch.ones(1, 3), torch.full((1, 3), 90.0), torch.tensor(n))

    def formula(z):          # what reduced_formula / the element tuple are keyed on
        c = Counter(z)
        g = 0
        for v in c.values():
            g = math.gcd(g, v)
        return tuple(sorted((e, v // g) for e, v in c.items())), tuple(sorted(c))

    rng = np.random.default_rng(3)
    mine, ref = LongTimeMem(device="cuda"), LongTimeMemOracle()
    pool = [[8, 8, 22], [22, 8, 8, 8, 8, 22], [26, 8], [8, 26, 26, 8], [3, 15, 8, 8, 8, 8], [8, 22], [22, 8, 8, 8]]
    for step in range(6):
        zs = [pool[i] for i in rng.integers(0, len(pool), 40)]
        rew = rng.random(40)
        mine.extend([crystal(z) for z in zs], rew, step)
        f = [formula(z) for z in zs]
        ref.extend([a for a, _ in f], [b for _, b in f], rew, step)
        for method in ("composition", "element_comb"):
            got = mine.div_filter([crystal(z) for z in pool], np.ones(len(pool)), tol=10, buff=20, method=method)
            want = ref.div_filter([formula(z)[0 if method == "composition" else 1] for z in pool], np.ones(len(pool)),
                                  tol=10, buff=20, method=method)
            assert np.allclose(got[0], want[0]) and got[1] == want[1] and got[2:] == want[2:], (step, method)
        assert mine.unique_comps.numel() == len(ref.unique_comps)
    assert abs(mine.get_baseline(5) - ref.get_baseline(5)) < 1e-12
