"""MatInvent RL loop — mirror of pipeline/mat_invent.py:17-290 with the fine-tune inner loop rewritten on the
device (`ft_step`, :125-189 -> models/diffcsp/finetune.FineTuner) and one-process-per-GPU data parallelism:
every rank runs the same loop with the same seeds, samples its shard of the crystal batch (no collective),
all-gathers the sampled crystals for scoring, and fine-tunes with one gradient all-reduce per Adam step.

The long-term memory and its diversity filter run on the device (memory/ltm.py, `ltm=True`).  Validity / SUN
filters and extxyz dumps are the reference's host-side subsystems (out of scope): hooks are called
when such objects are supplied."""
import logging
import os
import time

import numpy as np

from ..models.diffcsp.finetune import FineTuner
from .base import ReinL


def _dist():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist, dist.get_rank(), dist.get_world_size()
    return None, 0, 1


class MatInvent(ReinL):
    def __init__(self, rl_epoch, model_suite, reward, sample_cfg, finetune_cfg, save_dir, save_freq, device=None,
                 logger=None, replay=False, replay_args=None, topk_ratio=0.5, div_filter=False, df_args=None,
                 noise=None, **kwargs):
        super().__init__(rl_epoch=rl_epoch, model_suite=model_suite, reward=reward, sample_cfg=sample_cfg,
                         finetune_cfg=finetune_cfg, save_dir=save_dir, save_freq=save_freq, device=device,
                         logger=logger, replay=replay, replay_args=replay_args, **kwargs)
        self.topk_ratio = topk_ratio
        assert 0 < topk_ratio <= 1
        self.div_filter = bool(div_filter) and self.ltm is not None
        self.df_args = dict(df_args or {})
        self.noise = noise
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
    def sample_step(self):
        dist, rank, world = _dist()
        cfg = dict(self.sample_cfg)
        bs, nb = int(cfg.pop("batch_size")), int(cfg.pop("num_batches"))
        if world > 1:
            # shard the crystal batch: rank r samples ceil/floor share; no communication during the 1000 steps
            share = [bs // world + (1 if r < bs % world else 0) for r in range(world)]
            data, strucs = self.sampler.generate(self.agent, batch_size=max(share[rank], 1), num_batches=nb,
                                                 noise=self.noise, **cfg)
            gathered = [None] * world
            dist.all_gather_object(gathered, (data[:share[rank] * nb], strucs[:share[rank] * nb]))
            data = [d for part in gathered for d in part[0]]
            strucs = [s for part in gathered for s in part[1]]
        else:
            data, strucs = self.sampler.generate(self.agent, batch_size=bs, num_batches=nb, noise=self.noise, **cfg)
        return data, strucs, None, {}

    # ------------------------------------------------------------------ fine-tuning (:125-189)
    def ft_step(self, data_list, rewards, baseline=None):
        cfg = self.finetune_cfg
        dist, rank, world = _dist()
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
        sample_list, sample_struc, xyz_path, _ = self.sample_step()
        sample_list, sample_struc, rewards, prop_dict = self.reward_step(sample_list, sample_struc, xyz_path,
                                                                          "step_%04d" % self.step)
        log = {"reward mean": float(rewards.mean()), "reward std": float(rewards.std()), "cost": self.cost}
        penalty_strucs = []
        if self.ltm is not None:          # pipeline/mat_invent.py:209-237
            self.ltm.extend(sample_struc, rewards, self.step)
            if hasattr(self.ltm, "calc_metrics"):
                burden, div_ratio = self.ltm.calc_metrics(getattr(self.reward, "threshold", 0.0))
                log.update({"crystal_num": len(self.ltm), "unique_comps": len(self.ltm.unique_comps), "burden": burden,
                            "div_ratio": div_ratio})
            if self.div_filter:
                rewards, penalty_idx, tol_n, buff_n = self.ltm.div_filter(sample_struc, rewards, **self.df_args)
                penalty_strucs = [sample_struc[p] for p in penalty_idx]
                logging.info("Diversity filter: tol_n=%d, buff_n=%d", tol_n, buff_n)
        if self.logger is not None:
            self.logger.log(log, step=self.step)
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
        else:
            ft_data, ft_reward = sample_topk, reward_topk
        logs = self.ft_step(ft_data, ft_reward, None)
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
