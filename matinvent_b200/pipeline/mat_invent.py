"""MatInvent RL loop — mirror of pipeline/mat_invent.py:17-290 with the fine-tune inner loop rewritten on the
device (`ft_step`, :125-189 -> models/diffcsp/finetune.FineTuner) and one-process-per-GPU data parallelism:
every rank runs the same loop with the same seeds, samples its shard of the crystal batch (no collective),
all-gathers the sampled crystals for scoring, and fine-tunes with one gradient all-reduce per Adam step.

The long-term memory and its diversity filter run on the device (memory/ltm.py), so does the validity pre-filter
(pipeline/filters.py); sampled structures are written as extxyz (pipeline/utils.py).  MLIP relaxation and the SUN
filter are the reference's external subsystems (out of scope): `sample_cfg.mlip_opt` / `sample_cfg.filter`
are called when supplied, exactly where the reference calls them."""
import logging
import os
import time

import numpy as np

from ..models.diffcsp.finetune import FineTuner
from .base import ReinL
from .filters import invalid_filter
from .utils import save_structures


def _dist():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist, dist.get_rank(), dist.get_world_size()
    return None, 0, 1


class MatInvent(ReinL):
    def __init__(self, rl_epoch, model_suite, reward, sample_cfg, finetune_cfg, save_dir, save_freq, device=None,
                 logger=None, replay=False, replay_args=None, topk_ratio=0.5, div_filter=False, df_args=None,
                 noise=None, save_samples=True, **kwargs):
        super().__init__(rl_epoch=rl_epoch, model_suite=model_suite, reward=reward, sample_cfg=sample_cfg,
                         finetune_cfg=finetune_cfg, save_dir=save_dir, save_freq=save_freq, device=device,
                         logger=logger, replay=replay, replay_args=replay_args, **kwargs)
        self.topk_ratio = topk_ratio
        assert 0 < topk_ratio <= 1
        self.div_filter = bool(div_filter) and self.ltm is not None
        self.df_args = dict(df_args or {})
        self.noise = noise
        self.save_samples = bool(save_samples)
        self.timing = {}
        self.load_model()

    def load_model(self):
        """agent + frozen prior from the same checkpoint (pipeline/mat_invent.py:62-72)."""
        self.agent = self.model_suite.load_model().to(self.device)
        self.prior = self.model_suite.load_model().to(self.device)
        for p in self.agent.parameters():
            p.requires_grad = True
        for p in self.prior.parameters():
            p.requires_grad = False

    # ------------------------------------------------------------------ sampling (:74-123)
    NON_SAMPLER_KEYS = ("filter", "max_num", "mlip_opt", "invalid_filter", "smact_validity", "structure_validity")

    def sample_step(self):
        """generate -> invalid_filter -> (save) -> mlip_opt hook -> filter hook -> max_num truncation -> (save)
        (pipeline/mat_invent.py:74-123).

        Multi-rank (one process per GPU): every rank draws the GLOBAL atom-count list from the rank-identical numpy
        stream (exactly the draw a single process makes), samples its contiguous share balanced by sum n^2 with a
        rank-specific noise stream, and all ranks gather the post-processed crystals — no collective during the 1000
        reverse steps.  Everything after the gather is rank-identical."""
        dist, rank, world = _dist()
        cfg = dict(self.sample_cfg)
        bs, nb = int(cfg.pop("batch_size")), int(cfg.pop("num_batches"))
        hooks = {k: cfg.pop(k, None) for k in self.NON_SAMPLER_KEYS}
        t0 = time.time()
        if world > 1:
            from ..models.diffcsp.diffusion import PhiloxNoise
            from ..models.diffcsp.finetune import partition_crystals
            from ..models.diffcsp.sample import SampleDataset, pack_crystals, to_structure, unpack_crystals
            import torch
            counts = SampleDataset(bs * nb, self.sampler.num_atoms_distribution).num_atoms.tolist()
            data, strucs = [], []
            # rank-identical seed from torch's (rank-identical) CPU generator; rank-specific Philox stream from it
            base = int(torch.randint(0, 2 ** 31 - 1, (1,)))
            noise = self.noise if self.noise is not None else PhiloxNoise(self.device, seed=base * world + rank)
            for b in range(nb):
                chunk = [max(int(n), 1) for n in counts[b * bs:(b + 1) * bs]]
                lo, hi = partition_crystals(chunk, world)[rank]
                if hi > lo:
                    d, _ = self.sampler.generate(self.agent, batch_size=hi - lo, num_batches=1, noise=noise,
                                                 num_atoms=chunk[lo:hi], **cfg)
                else:
                    d = []
                # the shard travels as five concatenated tensors (one small pickle per rank, not five tensors per crystal)
                gathered = [None] * world
                dist.all_gather_object(gathered, pack_crystals(d))
                part = [x for g_ in gathered for x in unpack_crystals(g_)]
                data += part
                strucs += [to_structure(x) for x in part]
        else:
            data, strucs = self.sampler.generate(self.agent, batch_size=bs, num_batches=nb, noise=self.noise, **cfg)
        self.timing["sample_s"] = time.time() - t0
        t0 = time.time()
        n_gen = len(data)
        if callable(hooks["invalid_filter"]):             # the caller's own filter in place of the default
            data, strucs = hooks["invalid_filter"](data, strucs)
        elif hooks["invalid_filter"] is not False:
            data, strucs = invalid_filter(data, strucs, device=self.device,
                                          structure_validity=hooks["structure_validity"] if hooks["structure_validity"] is not None else True,
                                          smact_validity=hooks["smact_validity"])
        logging.info("Number of valid samples: %d of %d", len(strucs), n_gen)
        valid_xyz_path = None
        if self.save_samples and rank == 0:
            valid_xyz_path = save_structures(strucs, self.sample_dir, "step_%04d_valid.extxyz" % self.step)
        energies = None
        if hooks["mlip_opt"]:
            strucs, energies = hooks["mlip_opt"](strucs, valid_xyz_path)
        metrics = {}
        if hooks["filter"]:
            data, strucs, metrics = hooks["filter"](data, strucs, energies)
            logging.info("Number of filtered samples: %d", len(strucs))
        if hooks["max_num"] and len(strucs) > hooks["max_num"]:
            data, strucs = data[:hooks["max_num"]], strucs[:hooks["max_num"]]
        eval_xyz_path = None
        if self.save_samples and rank == 0:
            eval_xyz_path = save_structures(strucs, self.sample_dir, "step_%04d_eval.extxyz" % self.step)
        self.timing["filter_s"] = time.time() - t0
        return data, strucs, eval_xyz_path, metrics

    # ------------------------------------------------------------------ fine-tuning (:125-189)
    def ft_step(self, data_list, rewards, baseline=None):
        cfg = self.finetune_cfg
        dist, rank, world = _dist()
        if world > 1:
            # the shuffle, the replay picks and the global noise all assume rank-identical inputs and RNG streams
            import torch
            sig = torch.tensor([float(len(data_list)), float(sum(int(d.num_atoms) for d in data_list)),
                                float(np.asarray(rewards, dtype=np.float64).sum()), float(torch.rand(1))],
                               dtype=torch.float64, device=self.device)
            lo_, hi_ = sig.clone(), sig.clone()
            dist.all_reduce(lo_, op=dist.ReduceOp.MIN), dist.all_reduce(hi_, op=dist.ReduceOp.MAX)
            if not torch.equal(lo_, hi_):
                raise RuntimeError("fine-tune inputs or RNG streams differ between ranks (seed every rank identically): %s vs %s"
                                   % (lo_.tolist(), hi_.tolist()))
        loader = self.model_suite.get_dataloader(samples=data_list, rewards=rewards, batch_size=len(data_list))
        tuner = FineTuner(self.agent, self.prior, lr=cfg.lr, accum_steps=cfg.accum_steps, sigma=cfg.sigma,
                          rank=rank, world=world, noise=self.noise)
        logs = []
        for epoch in range(cfg.epochs):
            self.agent.train()
            loss_all = diff_all = kl_all = 0.0
            for batch in loader:
                loss, loss_diff, loss_kl = tuner.run_batch(batch, cfg.timesteps)
                loss_all += loss * batch.num_graphs
                diff_all += loss_diff
                kl_all += loss_kl
            n = len(data_list)
            logs.append({"loss": loss_all / n, "loss_diff": diff_all / n, "loss_kl": kl_all / n})
            logging.info("Epoch %d: %s", epoch, ", ".join("%s: %.4f" % kv for kv in logs[-1].items()))
        return logs

    # ------------------------------------------------------------------ one RL iteration (:191-271)
    def rl_step(self):
        t0 = time.time()
        self.timing = {}
        sample_list, sample_struc, xyz_path, sample_metrics = self.sample_step()
        t1 = time.time()
        sample_list, sample_struc, rewards, prop_dict = self.reward_step(sample_list, sample_struc, xyz_path,
                                                                          "step_%04d" % self.step)
        self.timing["reward_s"] = time.time() - t1
        t1 = time.time()
        log = {"%s mean" % k: float(np.mean(v)) for k, v in prop_dict.items() if len(v)}
        log.update({"%s std" % k: float(np.std(v)) for k, v in prop_dict.items() if len(v)})
        log.update({"reward mean": float(rewards.mean()), "reward std": float(rewards.std())})
        log.update(sample_metrics)
        penalty_strucs = []
        _, rank, _ = _dist()
        if self.ltm is not None:          # pipeline/mat_invent.py:209-237
            self.ltm.extend(sample_struc, rewards, self.step)
            burden, div_ratio = self.ltm.calc_metrics(getattr(self.reward, "threshold", 0.0))
            if self.save_samples and rank == 0:
                self.ltm.save(os.path.join(self.sample_dir, "long_term_memory.csv"))
            log.update({"crystal_num": len(self.ltm), "unique_comps": len(self.ltm.unique_comps), "burden": burden,
                        "div_ratio": div_ratio})
        log["cost"] = self.cost
        if self.logger is not None:
            self.logger.log(log, step=self.step)
        if self.div_filter:
            rewards, penalty_idx, tol_n, buff_n = self.ltm.div_filter(sample_struc, rewards, **self.df_args)
            penalty_strucs = [sample_struc[p] for p in penalty_idx]
            logging.info("Diversity filter: tol_n=%d, buff_n=%d", tol_n, buff_n)
        order = np.argsort(rewards)[::-1]
        topk = order[: int(self.finetune_cfg.batch_size * self.topk_ratio)]
        sample_topk = [sample_list[i] for i in topk]
        strucs_topk = [sample_struc[i] for i in topk]
        reward_topk = rewards[topk]
        if self.replay is not None:      # purge -> sample -> extend (:247-256)
            if penalty_strucs:
                self.replay.memory_purge(penalty_strucs)
            data_replay, reward_replay = self.replay.sample()
            ft_data = sample_topk + data_replay
            ft_reward = np.concatenate((reward_topk, np.asarray(reward_replay, dtype=float)))
            self.replay.extend(sample_topk, strucs_topk, reward_topk)
            logging.info("replay buffer size=%d", len(self.replay))
        else:
            ft_data, ft_reward = sample_topk, reward_topk
        self.timing["memory_s"] = time.time() - t1
        t1 = time.time()
        baseline = None
        if self.ltm is not None and len(ft_reward):
            baseline = min(self.ltm.get_baseline(self.step), float(np.min(ft_reward)))      # :262-263 (unused by ft_step)
        logs = self.ft_step(ft_data, ft_reward, baseline) if len(ft_data) else []
        self.timing["finetune_s"] = time.time() - t1
        self.timing["total_s"] = time.time() - t0
        logging.info("LOOP %d finished in %.2f min", self.step, (time.time() - t0) / 60)
        return log, logs

    def run_rl(self):
        _, rank, _ = _dist()
        for step in range(self.rl_epoch):
            self.step = step
            self.rl_step()
            if (step + 1) % self.save_freq == 0 and rank == 0:
                self.model_suite.save_model(self.agent, os.path.join(self.models_dir, "loop_%04d" % step))
        if rank == 0:
            self.model_suite.save_model(self.agent, os.path.join(self.models_dir, "final"))
