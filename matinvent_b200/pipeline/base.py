"""RL pipeline base — mirror of pipeline/base.py:26-142 (`ReinL`): same constructor arguments, same config
merge (suite config overridden by pipeline config, :53-59), same `reward_step`.  The long-term memory is the
device-resident `memory.ltm.LongTimeMem` (created unconditionally like pipeline/base.py:64; `ltm=<object>` substitutes any
object with the same methods, `ltm=None` disables it); loggers and reward calculators of the reference are host-side
bookkeeping / external property oracles (out of scope): duck-typed — pass the reference's
objects or any stand-in."""
import logging
import os

import numpy as np

from ..config import Config
from ..memory.ltm import LongTimeMem
from ..memory.replay_buffer import ReplayBuffer
from ..models.suite.base import get_device


class ReinL:
    def __init__(self, rl_epoch, model_suite, reward, sample_cfg, finetune_cfg, save_dir, save_freq, device=None,
                 logger=None, replay=False, replay_args=None, ltm=True, **kwargs):
        self.rl_epoch = rl_epoch
        self.model_suite = model_suite
        self.reward = reward
        self.save_dir = save_dir
        self.save_freq = save_freq
        self.logger = logger
        self.device = get_device(device)
        self.cfg = Config(kwargs)
        self.step = 0
        self.cost = 0
        self.sample_cfg = Config.merge(model_suite.sample_cfg, sample_cfg)
        self.finetune_cfg = Config.merge(model_suite.finetune_cfg, finetune_cfg)
        self.sampler = model_suite.get_sampler()
        self.ltm = LongTimeMem(device=self.device) if ltm is True else ltm
        self.models_dir = os.path.join(save_dir, "models")
        self.sample_dir = os.path.join(save_dir, "samples")
        os.makedirs(self.models_dir, exist_ok=True)
        os.makedirs(self.sample_dir, exist_ok=True)
        self.replay = ReplayBuffer(device=self.device, **(replay_args or {})) if replay else None

    def reward_step(self, sample_data, sample_struc, xyz_path, label="tmp"):
        """pipeline/base.py:98-127: score, count the cost, drop failed samples."""
        rewards, prop_dict, failed_mask = self.reward.scoring((sample_struc, xyz_path), label)
        self.cost += len(sample_struc)
        failed_mask = np.asarray(failed_mask, dtype=bool)
        ok_rewards = np.asarray(rewards)[~failed_mask].astype(float)
        ok_props = {k: np.asarray(v)[~failed_mask] for k, v in prop_dict.items()}
        ok_data = [d for d, f in zip(sample_data, failed_mask) if not f]
        ok_struc = [s for s, f in zip(sample_struc, failed_mask) if not f]
        logging.info("Evaluation costs to date: %d", self.cost)
        logging.info("Number of samples that successfully obtained rewards: %d", len(ok_struc))
        if len(ok_rewards):
            logging.info("reward mean=%.4f std=%.4f", ok_rewards.mean(), ok_rewards.std())
        return ok_data, ok_struc, ok_rewards, ok_props

    def load_model(self):
        raise NotImplementedError

    def sample_step(self):
        raise NotImplementedError

    def ft_step(self, data_list):
        raise NotImplementedError

    def rl_step(self):
        raise NotImplementedError

    def run_rl(self):
        raise NotImplementedError
