"""torchrun entry (2+ ranks, one GPU each, NCCL): the sharded fine-tune step must reproduce the single-GPU
result (same global noise, loss scaled by 1/B_global, SUM all-reduce) and sharded sampling must return the
same crystals.  Launched by tests/test_gpu_pipeline.py::test_two_gpu_ft_and_sampling_match_single_gpu."""
import argparse
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--backend", default="nccl")
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group(args.backend, device_id=dev)
    from conftest import build_module, load_gold
    from test_gpu_pipeline import _ft_batch
    from matinvent_b200.models.diffcsp import PhiloxNoise
    from matinvent_b200.models.diffcsp.finetune import FineTuner
    gs = load_gold("small_net.pt")
    hp = gs["hp"]
    data, batch = _ft_batch([3, 9, 1, 14, 6, 20, 2], 4)

    def run(world_, rank_):
        agent = build_module(hp, gs["sd"], gs["sigmas_norm"], device=dev)
        prior = build_module(hp, gs["sd_prior"], gs["sigmas_norm"], device=dev)
        t = FineTuner(agent, prior, lr=1e-4, accum_steps=5, sigma=0.025, rank=rank_, world=world_,
                      noise=PhiloxNoise(dev, seed=11))
        logs = t.run_batch(batch, 20)
        return agent.decoder.flat.data.clone(), logs

    w_multi, logs_multi = run(world, rank)
    w_single, logs_single = run(1, 0)
    # replicas identical across ranks
    ws = [torch.empty_like(w_multi) for _ in range(world)]
    dist.all_gather(ws, w_multi)
    for w in ws[1:]:
        assert torch.equal(ws[0], w), "replicas diverged"
    d = (w_multi - w_single).abs()
    frac_big = float((d > 2e-6).float().mean())
    assert frac_big < 1e-3 and float(d.mean()) < 1e-7, (frac_big, float(d.mean()))
    for a, b in zip(logs_multi, logs_single):
        assert abs(a - b) < 1e-4 * abs(b) + 1e-9, (logs_multi, logs_single)
    # sharded sampling through the pipeline mirror: every rank samples its share, all ranks end up with the full list
    from matinvent_b200.pipeline.mat_invent import MatInvent
    import numpy as np

    class _Suite:
        sample_cfg, finetune_cfg = dict(batch_size=6, num_batches=1), dict(batch_size=4)
        def get_sampler(self):
            from matinvent_b200.models.diffcsp import DiffCSPSampler
            return DiffCSPSampler(batch_size=6, num_batches=1)
        def load_model(self):
            return build_module(hp, gs["sd"], gs["sigmas_norm"], device=dev)
    np.random.seed(100)             # every rank runs the loop with the SAME seeds (pipeline/mat_invent.py docstring): the
    torch.manual_seed(100)          # global atom-count draw, the fine-tune shuffle and the replay picks are rank-identical
    import tempfile
    pipe = MatInvent(rl_epoch=1, model_suite=_Suite(), reward=None, sample_cfg=dict(invalid_filter=False), finetune_cfg={},
                     save_dir=tempfile.mkdtemp(),
                     save_freq=1, device=str(dev), save_samples=False)
    data, strucs, _, _ = pipe.sample_step()
    assert len(data) == 6 and len(strucs) == 6, len(data)
    sizes = torch.tensor([int(d.num_atoms) for d in data], device=dev)
    all_sizes = [torch.empty_like(sizes) for _ in range(world)]
    dist.all_gather(all_sizes, sizes)
    for t in all_sizes[1:]:
        assert torch.equal(all_sizes[0], t), "ranks disagree on the gathered sample list"
    # the union over ranks is the draw a single process makes from the same numpy stream, and ranks sampled with
    # different noise streams (identical crystals would mean a shared stream)
    np.random.seed(100)
    from matinvent_b200.models.diffcsp.sample import SampleDataset
    assert [max(int(n), 1) for n in SampleDataset(6).num_atoms.tolist()] == sizes.tolist()
    fr = [d.frac_coords for d in data]
    assert not any(a.shape == b.shape and torch.equal(a, b) for i, a in enumerate(fr) for b in fr[i + 1:])
    if rank == 0:
        print("DIST_CHECK_OK", logs_multi)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
