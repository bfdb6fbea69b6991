"""One RL iteration of BASELINE configs[4] shape ("multi-objective RL (hhi + magmom) with device replay_buffer, 10k
crystals/iter, 8xB200") through matinvent_b200.pipeline.MatInvent, with the wall-time breakdown of its stages.

    python scripts/run_rl_iteration.py --crystals 512                                   # 1 GPU, plumbing
    torchrun --nproc-per-node 8 scripts/run_rl_iteration.py --crystals 10240            # the configuration itself

Full-size random-init DiffCSP (the benchmark network), the complete 1000-step sampler, device validity pre-filter (loose
thresholds: a random-init net does not produce physical cells), device composition reward (min of two scaled table
properties; synthetic element tables — pymatgen's are not in this image), long-term memory + diversity filter, device
replay buffer, reward-weighted fine-tune epoch with one gradient all-reduce per Adam step.  Rank 0 prints one JSON line."""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--crystals", type=int, default=10240, help="crystals sampled per RL iteration (global)")
    ap.add_argument("--per-gpu-batch", type=int, default=256)
    ap.add_argument("--ft-batch", type=int, default=64, help="fine-tune batch: top-k half + replay half")
    ap.add_argument("--ft-timesteps", type=int, default=1000)
    ap.add_argument("--iterations", type=int, default=2, help="the second iteration runs with warm graphs and a filled replay buffer")
    args = ap.parse_args()
    import torch.distributed as dist
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from matinvent_b200.models.suite import DiffCSPSuite
    from matinvent_b200.pipeline import MatInvent
    from matinvent_b200.pipeline.filters import invalid_filter
    from matinvent_b200.rewards import CompositionReward, synthetic_table
    np.random.seed(0)
    torch.manual_seed(0)
    gbatch = args.per_gpu_batch * world
    nb = max(1, args.crystals // gbatch)
    sig = torch.load(os.path.join(ROOT, "tests", "golden", "sigmas_norm_T1000.pt"))["sigmas_norm"]
    suite = DiffCSPSuite(model_name="diffcsp", sample_cfg=dict(batch_size=gbatch, num_batches=nb),
                         finetune_cfg=dict(batch_size=args.ft_batch, timesteps=args.ft_timesteps, lr=1e-4), device=str(dev),
                         random_init=True, head_scale=0.05, seed=0)
    reward = CompositionReward(
        prop_cfg=[dict(name="hhi", table=synthetic_table("hhi"), target="descending", minv=750, maxv=3250),
                  dict(name="magmom", table=synthetic_table("magmom"), weights="atom", target="ascending", minv=0.0, maxv=0.25)],
        reward_threshold=0.8, reduce="min", device=str(dev))
    save_dir = tempfile.mkdtemp(prefix="mi_rl_")
    pipe = MatInvent(rl_epoch=args.iterations, model_suite=suite, reward=reward,
                     sample_cfg=dict(invalid_filter=lambda d, s: invalid_filter(d, s, device=str(dev), structure_validity=False,
                                                                                max_len=1e30)),
                     finetune_cfg=dict(batch_size=args.ft_batch, accum_steps=50, epochs=1, sigma=0.025, timesteps=args.ft_timesteps),
                     save_dir=save_dir, save_freq=10 ** 6, device=str(dev), replay=True,
                     replay_args=dict(buffer_size=1000, sample_size=args.ft_batch // 2, reward_cutoff=0.0),
                     div_filter=True, df_args=dict(tol=5, buff=10), save_samples=(rank == 0))
    for m in (pipe.agent, pipe.prior):
        m.sigma_scheduler.sigmas_norm.copy_(sig.to(m.sigma_scheduler.sigmas_norm.device))
    out = []
    for it in range(args.iterations):
        pipe.step = it
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.time()
        log, logs = pipe.rl_step()
        torch.cuda.synchronize()
        wall = time.time() - t0
        t = dict(pipe.timing)
        out.append(dict(iteration=it, wall_s=wall, sample_s=t.get("sample_s"), filter_s=t.get("filter_s"), reward_s=t.get("reward_s"),
                        memory_s=t.get("memory_s"), finetune_s=t.get("finetune_s"), scored=int(pipe.cost), replay=len(pipe.replay),
                        ltm=len(pipe.ltm), reward_mean=log.get("reward mean"), loss=logs[0]["loss"] if logs else None))
    if rank == 0:
        print(json.dumps(dict(config="one RL iteration: %d crystals sampled (%d x %d per call) on %d GPU(s), full 1000-step sampler, "
                                     "multi-objective composition reward, diversity filter, device replay buffer, fine-tune batch %d "
                                     "x %d timesteps (accum 50)" % (gbatch * nb, nb, gbatch, world, args.ft_batch, args.ft_timesteps),
                              n_gpus=world, crystals_per_iteration=gbatch * nb, iterations=out,
                              crystals_per_s_sampling=gbatch * nb / out[-1]["sample_s"])))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
