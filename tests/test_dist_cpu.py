"""CPU, world_size 2, gloo: the host-side sharding logic of the fine-tune step.

The CUDA engine cannot run here, so each rank computes ITS shard's gradient with the oracle (torch autograd on
the CPU) exactly the way FineTuner does — contiguous slice from partition_crystals, noise drawn for the GLOBAL
batch and sliced, loss scaled by 1/(B_global*accum) — then SUM-all-reduces; the result must equal the
single-process gradient of the global batch."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, load_gold


def test_partition_crystals_contiguous_and_balanced():
    from matinvent_b200.models.diffcsp.finetune import partition_crystals
    na = [3, 9, 1, 14, 6, 20, 2, 8, 8, 8]
    for world in (1, 2, 3, 4):
        parts = partition_crystals(na, world)
        assert len(parts) == world and parts[0][0] == 0 and parts[-1][1] == len(na)
        for (a, b), (c, d) in zip(parts[:-1], parts[1:]):
            assert b == c and a <= b
        loads = [sum(n * n for n in na[a:b]) for a, b in parts]
        assert max(loads) <= sum(loads) / world + max(n * n for n in na)
    assert partition_crystals([5], 1) == [(0, 1)]


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sys.path.insert(0, ROOT)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import diffcsp_oracle as O
        from oracle.ref_import import make_batch
        from matinvent_b200.models.diffcsp.finetune import partition_crystals
        torch.set_num_threads(2)
        gs = load_gold("small_net.pt")
        hp, sch = gs["hp"], O.Schedules(gs["hp"], gs["sigmas_norm"])
        na = [3, 9, 1, 14, 6]
        g = torch.Generator().manual_seed(4)
        N, B = sum(na), len(na)
        cr = dict(lengths=3 + 5 * torch.rand(B, 3, generator=g), angles=70 + 40 * torch.rand(B, 3, generator=g),
                  frac_coords=torch.rand(N, 3, generator=g), atom_types=torch.randint(1, 101, (N,), generator=g))
        reward = torch.rand(B, generator=g)
        accum, sigma, t_idx = 50, 0.025, 7

        class SlicedNoise:
            """global draw, local slice — what FineTuner does with its z_l/z_x/z_a buffers"""
            def __init__(self, lo, hi, nlo, nhi):
                self.gen = torch.Generator().manual_seed(99)
                self.lo, self.hi, self.nlo, self.nhi = lo, hi, nlo, nhi
            def randn(self, shape):
                shape = tuple(shape)
                full = {(self.hi - self.lo, 3, 3): (B, 3, 3), (self.nhi - self.nlo, 3): (N, 3),
                        (self.nhi - self.nlo, 100): (N, 100)}[shape]
                t = torch.randn(full, generator=self.gen)
                return t[self.lo:self.hi] if full[0] == B and len(full) == 3 else t[self.nlo:self.nhi]

        def grad_of(lo, hi, scale_B):
            nlo, nhi = sum(na[:lo]), sum(na[:hi])
            sd = {k: v.clone().requires_grad_(True) for k, v in gs["sd"].items()}
            b = make_batch(na[lo:hi], lengths=cr["lengths"][lo:hi], angles=cr["angles"][lo:hi],
                           frac_coords=cr["frac_coords"][nlo:nhi], atom_types=cr["atom_types"][nlo:nhi])
            loss, _ = O.ft_timestep_loss(sd, gs["sd_prior"], hp, sch, b, reward[lo:hi], t_idx,
                                         SlicedNoise(lo, hi, nlo, nhi), sigma, accum)
            (loss * (hi - lo) / scale_B).backward()          # mean over local -> sum/(B_global)
            return torch.cat([v.grad.reshape(-1) for v in sd.values()])

        lo, hi = partition_crystals(na, world)[rank]
        gl = grad_of(lo, hi, B)
        dist.all_reduce(gl)
        if rank == 0:
            ref = grad_of(0, B, B)
            ret["err"] = float((gl - ref).abs().max() / ref.abs().max())
    finally:
        dist.destroy_process_group()


def test_sharded_gradient_equals_global_gradient_gloo():
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, 2, 29533, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert ret["err"] < 1e-5, ret["err"]


def test_fine_tune_group_size_divides_the_accumulation_window():
    """host logic of the stacked-timestep fine-tune step: the group is a divisor of accum_steps (a group never
    straddles an optimizer step) bounded by the edge budget; an explicit group is clamped the same way"""
    import types
    from matinvent_b200.models.diffcsp.finetune import FineTuner
    f = lambda accum, edges, group=None: FineTuner.group_size(
        types.SimpleNamespace(accum=accum, group=group, GROUP_EDGES=FineTuner.GROUP_EDGES), edges)
    assert f(50, 2786) == 50                       # the reference's working point: 18 crystals, whole window in one launch
    assert f(50, 40000) == 2 and f(50, 200000) == 1
    assert f(50, 2786, group=10) == 10 and f(50, 2786, group=7) == 5 and f(50, 2786, group=1000) == 50
    assert f(7, 100) == 7 and f(7, 30000) == 1     # prime window: all or one
    for accum in (1, 10, 48, 50):
        for edges in (1, 500, 5000, 60000):
            g = f(accum, edges)
            assert accum % g == 0 and (g == 1 or g * edges <= FineTuner.GROUP_EDGES)
