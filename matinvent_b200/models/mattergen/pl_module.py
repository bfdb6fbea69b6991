"""MatterGen back-end adapter — mirror of models/mattergen/pl_module.py:16-102 and models/mattergen/loss.py:11-78.

What the reference keeps IN ITS TREE for this back-end is the adapter between `MatInvent.ft_step`
(pipeline/mat_invent.py:150-164) and the `mattergen` package: the fine-tune time grid (`add_noise`), the per-sample
weighted loss (`SampleLoss`) and the KL proxy (`calc_kl_reg`).  Those are implemented here on the device and pinned to
the unmodified reference files (oracle/mattergen_oracle.py, tests/test_oracle_vs_reference.py).  The score network
(GemNet-T), the corruption processes and the per-field denoising losses live in microsoft/mattergen@5bb2b39, which
is neither vendored in the reference nor installed here: like the reference, this module takes them as a
`diffusion_module` object (duck-typed: `.model(noisy, t)`, `.corruption.T`, `.corruption.sample_marginal(batch, t)`,
`.corruption.corruptions`, `.pre_corruption_fn(batch)`, `._get_device(batch)`, `.loss_fn.loss_fns`) — the real
package's objects drop in unchanged where it is installed."""
import torch
import torch.nn as nn

from ... import ops

DEFAULT_WEIGHTS = {"atomic_numbers": 1.0, "cell": 1.0, "pos": 0.1}          # loss.py:21-26


class _WeightedFieldSum(torch.autograd.Function):
    """agg[b] = sum_k w_k v_k[b] on the device (mi_weighted_field_sum); d agg / d v_k = w_k"""

    @staticmethod
    def forward(ctx, weights, *fields):
        vals = [v.detach().to(torch.float32).contiguous() for v in fields]
        out = torch.empty_like(vals[0])
        ops.weighted_field_sum(vals, weights, out)
        ctx.weights = weights
        return out

    @staticmethod
    def backward(ctx, gout):
        return (None,) + tuple(gout * w for w in ctx.weights)


class SampleLoss:
    """loss.py:11-78: per-field per-sample losses from the injected loss functions, aggregated PER SAMPLE with the weights
    pos 0.1 / cell 1.0 / atomic_numbers 1.0 (the reference's MaterialsLoss subclass returns [batch_size], not a scalar)."""

    def __init__(self, loss_fns, reduce="sum", d3pm_hybrid_lambda=0.01, weights=None):
        self.loss_fns = dict(loss_fns)
        self.reduce, self.d3pm_hybrid_lambda = reduce, d3pm_hybrid_lambda
        self.loss_weights = dict(weights or DEFAULT_WEIGHTS)

    def __call__(self, *, multi_corruption, batch, noisy_batch, score_model_output, t, node_is_unmasked=None):
        B = batch.get_batch_size()
        per_field = {}
        for k, fn in self.loss_fns.items():      # mattergen's `apply`: every keyword indexed by field, plus the broadcast ones
            per_field[k] = fn(corruption=multi_corruption.corruptions[k], x=batch[k], noisy_x=noisy_batch[k],
                              score_model_output=score_model_output[k], batch_idx=batch.get_batch_idx(k),
                              node_is_unmasked=node_is_unmasked, t=t, batch_size=B, batch=batch)
        assert set(tuple(v.shape) for v in per_field.values()) == {(B,)}, "All losses should have shape (batch_size,)."
        metrics = {k: v.mean() for k, v in per_field.items()}
        keys = list(per_field)
        if len(keys) > 4:
            raise NotImplementedError("at most four loss fields")
        agg = _WeightedFieldSum.apply([self.loss_weights[k] for k in keys], *[per_field[k] for k in keys])
        return agg, metrics


class _KLProxy(torch.autograd.Function):
    """pl_module.py:83-102 via mi_rl_loss (the same three per-crystal mean squared differences as the DiffCSP back-end's
    calc_kl_reg, diffusion.py:140-149), with the analytic gradient w.r.t. the agent's predictions"""

    @staticmethod
    def forward(ctx, node_off, B, cell, pos, logits, cell_p, pos_p, logits_p):
        c = lambda v: v.detach().to(torch.float32).contiguous()
        pred, prior = (c(cell), c(pos), c(logits)), (c(cell_p), c(pos_p), c(logits_p))
        kl = torch.empty(B, device=pos.device)
        ops.rl_loss(pred, None, prior, node_off, B, logits.shape[1], (1.0, 1.0, 1.0), None, None, 1.0, None, kl, None)
        ctx.args = (node_off, B, pred, prior)
        return kl

    @staticmethod
    def backward(ctx, gout):
        node_off, B, pred, prior = ctx.args
        d = tuple(torch.empty_like(v) for v in pred)
        ops.rl_loss(pred, None, prior, node_off, B, pred[2].shape[1], (1.0, 1.0, 1.0), None, gout.contiguous(), 1.0, None, None, d)
        return (None, None, d[0], d[1], d[2], None, None, None)


class MatterGenModule(nn.Module):
    N_FT = 1000          # pl_module.py:59

    def __init__(self, diffusion_module, optimizer_partial=None, scheduler_partials=None):
        super().__init__()
        self.diffusion_module = diffusion_module
        self.optimizer_partial, self.scheduler_partials = optimizer_partial, scheduler_partials
        self.sample_loss_fn = SampleLoss(loss_fns=diffusion_module.loss_fn.loss_fns)

    # ------------------------------------------------------------------ pl_module.py:55-69
    def add_noise(self, batch, timestep):
        batch = self.diffusion_module.pre_corruption_fn(batch)
        max_t = self.diffusion_module.corruption.T
        device = self.diffusion_module._get_device(batch)
        time_list = torch.linspace(max_t, 1 / self.N_FT, self.N_FT, device=device)      # a 1000-entry table, built like the reference's
        t = torch.full((batch.get_batch_size(),), time_list[timestep], device=device)
        noisy_batch = self.diffusion_module.corruption.sample_marginal(batch, t)
        return noisy_batch, batch, t

    # ------------------------------------------------------------------ pl_module.py:71-81
    def calc_sample_loss(self, noised_input):
        noisy_batch, batch, t = noised_input
        score_model_output = self.diffusion_module.model(noisy_batch, t)
        loss, _ = self.sample_loss_fn(multi_corruption=self.diffusion_module.corruption, batch=batch, noisy_batch=noisy_batch,
                                      score_model_output=score_model_output, t=t)
        return loss, score_model_output

    # ------------------------------------------------------------------ pl_module.py:83-102
    def calc_kl_reg(self, agent_pred, prior_pred, batch):
        batch_idx = batch.get_batch_idx("pos")
        B = int(batch.get_batch_size())
        counts = torch.bincount(batch_idx, minlength=B)
        node_off = torch.zeros(B + 1, dtype=torch.int32, device=batch_idx.device)
        node_off[1:] = torch.cumsum(counts, 0).to(torch.int32)
        return _KLProxy.apply(node_off, B, agent_pred["cell"], agent_pred["pos"], agent_pred["atomic_numbers"],
                              prior_pred["cell"].detach(), prior_pred["pos"].detach(), prior_pred["atomic_numbers"].detach())
