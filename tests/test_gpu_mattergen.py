"""GPU: the MatterGen adapter — the arithmetic the reference keeps in its tree, on the device,
against oracle/mattergen_oracle.py (pinned to the unmodified models/mattergen/{pl_module,loss}.py).  The mattergen
package's own leaves (score network, corruptions, per-field losses, predictor-corrector) are injected stand-ins."""
import types

import numpy as np
import pytest
import torch

from conftest import rel_err
from test_oracle_vs_reference import _FakeBatch

pytestmark = pytest.mark.gpu


def _setup(seed=0):
    g = torch.Generator().manual_seed(seed)
    na = torch.tensor([3, 1, 7, 20, 4, 11])
    B, N = len(na), int(na.sum())
    bi = torch.repeat_interleave(torch.arange(B), na).cuda()
    fields = {"pos": torch.rand(N, 3, generator=g).cuda(), "cell": torch.randn(B, 3, 3, generator=g).cuda(),
              "atomic_numbers": torch.randint(1, 101, (N,), generator=g).cuda()}
    return g, na, B, N, bi, _FakeBatch(fields, bi, B)


class _StubScore(torch.nn.Module):
    """stand-in score model: a linear read-out of the noisy fields (GemNet-T is un-vendored)"""

    def __init__(self, A=101):
        super().__init__()
        self.w = torch.nn.Parameter(torch.tensor([0.7, -0.3, 1.1]))
        self.A = A

    def forward(self, noisy, t):
        pos, cell = noisy["pos"], noisy["cell"]
        logits = (pos.sum(1, keepdim=True) * self.w[2] + t[noisy.get_batch_idx("pos")][:, None]).expand(-1, self.A) * \
            torch.linspace(-1, 1, self.A, device=pos.device)
        return {"pos": pos * self.w[0], "cell": cell * self.w[1], "atomic_numbers": logits.contiguous()}


def _diffusion_module(model, per_field):
    class Corr:
        T = 1.0
        corruptions = {"pos": "cp", "cell": "cc", "atomic_numbers": "ca"}

        def sample_marginal(self, batch, t):
            f = dict(batch.fields)
            f["pos"] = (f["pos"] + t[batch.get_batch_idx("pos")][:, None]) % 1.0
            return _FakeBatch(f, batch._bi, batch.B)

    loss_fns = {k: (lambda k_: (lambda **kw: per_field(k_, **kw)))(k) for k in ("pos", "cell", "atomic_numbers")}
    return types.SimpleNamespace(pre_corruption_fn=lambda b: b, corruption=Corr(), _get_device=lambda b: torch.device("cuda"),
                                 model=model, loss_fn=types.SimpleNamespace(loss_fns=loss_fns))


def test_mattergen_module_matches_oracle():
    from oracle import mattergen_oracle as MO
    from matinvent_b200.models.mattergen import MatterGenModule
    g, na, B, N, bi, batch = _setup()

    def per_field(k, score_model_output, batch_idx, batch_size, **kw):          # a differentiable per-sample loss per field
        v = score_model_output.reshape(score_model_output.shape[0], -1).pow(2).mean(1)
        if batch_idx is None:
            return v
        return torch.zeros(batch_size, device=v.device).index_add_(0, batch_idx, v)

    agent = MatterGenModule(_diffusion_module(_StubScore().cuda(), per_field))
    prior = MatterGenModule(_diffusion_module(_StubScore().cuda(), per_field))
    with torch.no_grad():
        prior.diffusion_module.model.w.mul_(0.9)
    for ts in (0, 500, 999):
        noisy, b2, t = agent.add_noise(batch, ts)
        assert torch.equal(t.cpu(), torch.full((B,), float(MO.finetune_time(1.0, ts, device="cuda"))))
    noised = agent.add_noise(batch, 321)
    loss, pred = agent.calc_sample_loss(noised)
    per = {k: per_field(k, score_model_output=pred[k], batch_idx=batch.get_batch_idx(k), batch_size=B) for k in ("pos", "cell", "atomic_numbers")}
    want, _ = MO.aggregate_sample_loss({k: v.detach().cpu() for k, v in per.items()})
    assert loss.shape == (B,) and rel_err(loss, want) < 1e-6
    with torch.no_grad():
        _, pred_p = prior.calc_sample_loss(noised)
    kl = agent.calc_kl_reg(pred, pred_p, batch)
    want_kl = MO.kl_reg({k: v.detach().cpu().double() for k, v in pred.items()}, {k: v.cpu().double() for k, v in pred_p.items()},
                        bi.cpu(), B)
    assert rel_err(kl, want_kl) < 1e-5
    # the reference's objective (pipeline/mat_invent.py:155-164) back-propagates through both device functions
    reward = torch.rand(B, generator=g).cuda()
    ((reward * loss + 0.025 * (1.1 - reward) * kl).mean() / 50).backward()
    gw = agent.diffusion_module.model.w.grad.clone()
    # same objective with plain torch autograd on the oracle's formulas
    m2 = _StubScore().cuda()
    pred2 = m2(noised[0], noised[2])
    per2 = {k: per_field(k, score_model_output=pred2[k], batch_idx=batch.get_batch_idx(k), batch_size=B) for k in ("pos", "cell", "atomic_numbers")}
    loss2 = sum(MO.DEFAULT_WEIGHTS[k] * per2[k] for k in per2)
    k0 = (pred2["cell"] - pred_p["cell"]).pow(2).mean(dim=(1, 2))
    sm = lambda v: torch.zeros(B, device="cuda").index_add_(0, bi, v) / torch.bincount(bi, minlength=B).clamp(min=1)
    kl2 = k0 + sm((pred2["pos"] - pred_p["pos"]).pow(2).mean(1)) + sm((pred2["atomic_numbers"] - pred_p["atomic_numbers"]).pow(2).mean(1))
    ((reward * loss2 + 0.025 * (1.1 - reward) * kl2).mean() / 50).backward()
    assert rel_err(gw, m2.w.grad) < 1e-5


def test_mattergen_sampler_returns_means_and_params():
    """draw_samples_from_sampler (sample.py:27-64): the MEAN of the sampler is kept, cells become (lengths, angles) on the
    device, crystals are split per num_atoms"""
    from oracle import diffcsp_oracle as O
    from matinvent_b200.models.mattergen import MatterGenSampler

    class StubPC:
        def sample(self, cond, mask):
            na = cond["num_atoms"]
            N = int(na.sum())
            g = torch.Generator().manual_seed(int(N))
            mean = dict(pos=torch.rand(N, 3, generator=g).cuda(), cell=(4 * torch.eye(3) + torch.randn(len(na), 3, 3, generator=g)).cuda(),
                        atomic_numbers=torch.randint(1, 101, (N,), generator=g).cuda(), num_atoms=na)
            sample = {k: (v + 1 if v.dtype.is_floating_point else v) for k, v in mean.items()}
            self.last = mean
            return sample, mean

    pc = StubPC()
    np.random.seed(0)
    model = torch.nn.Linear(1, 1).cuda()
    data, strucs = MatterGenSampler(batch_size=7, num_batches=2, sampler_factory=lambda m: pc).generate(model)
    assert len(data) == len(strucs) == 14
    last = pc.last
    lengths, angles = O.lattices_to_params_shape(last["cell"].cpu())
    off = [0] + torch.cumsum(last["num_atoms"].cpu(), 0).tolist()
    for i, d in enumerate(data[7:]):
        assert torch.equal(d.frac_coords, last["pos"].cpu()[off[i]:off[i + 1]])            # the mean, not the noisy sample
        assert torch.equal(d.atom_types, last["atomic_numbers"].cpu()[off[i]:off[i + 1]])
        assert rel_err(d.lengths, lengths[i:i + 1]) < 1e-6 and float((d.angles - angles[i:i + 1]).abs().max()) < 1e-3
    with pytest.raises(RuntimeError):
        MatterGenSampler(batch_size=2, num_batches=1).generate(model)
