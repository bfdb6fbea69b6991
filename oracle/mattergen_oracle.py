"""TEST INFRASTRUCTURE — CPU restatement of the arithmetic the reference's MatterGen adapter holds IN THE TREE
(models/mattergen/pl_module.py:55-102, models/mattergen/loss.py:36-78).  Only tests/ may import this module.

PINNED: every function below is checked against the unmodified reference files imported under oracle/shims (stub
`mattergen` leaves) in tests/test_oracle_vs_reference.py.  What those files DELEGATE to the un-vendored package
microsoft/mattergen@5bb2b397a36de85a8dc9583b7d1d6353989de72c (env.yml:31) — the GemNet-T score network, the corruption
processes, the per-field loss functions, the predictor-corrector sampler — is not restated here: PARITY UNPINNED for
those, they are injected objects on both sides."""
import torch


def finetune_time(max_t, timestep, N=1000, device="cpu"):
    """pl_module.py:57-67: the fine-tune time of index `timestep`: linspace(max_t, 1/N, N)[timestep] (fp32, on `device`)"""
    return torch.linspace(max_t, 1 / N, N, device=device)[timestep]


DEFAULT_WEIGHTS = {"atomic_numbers": 1.0, "cell": 1.0, "pos": 0.1}          # loss.py:21-26


def aggregate_sample_loss(loss_per_sample_per_field, weights=None):
    """loss.py:63-73: per-sample weighted sum over fields (in the dict's order) + per-field batch means as metrics"""
    weights = weights or DEFAULT_WEIGHTS
    metrics = {k: v.mean() for k, v in loss_per_sample_per_field.items()}
    agg = torch.stack([weights[k] * v for k, v in loss_per_sample_per_field.items()], dim=0).sum(0)
    return agg, metrics


def kl_reg(agent_pred, prior_pred, batch_idx, num_graphs):
    """pl_module.py:83-102: mean squared differences of the cell [B,3,3], pos [N,3] and atomic-number logits [N,A]
    predictions of agent and (detached) prior, per crystal, unweighted sum"""
    def smean(v):
        out = torch.zeros(num_graphs, dtype=v.dtype).index_add_(0, batch_idx, v)
        cnt = torch.zeros(num_graphs, dtype=v.dtype).index_add_(0, batch_idx, torch.ones_like(v)).clamp_(min=1)
        return out / cnt
    k0 = torch.pow(agent_pred["cell"] - prior_pred["cell"].detach(), 2).mean(dim=(1, 2))
    k1 = smean(torch.pow(agent_pred["pos"] - prior_pred["pos"].detach(), 2).mean(dim=1))
    k2 = smean(torch.pow(agent_pred["atomic_numbers"] - prior_pred["atomic_numbers"].detach(), 2).mean(dim=1))
    return k0 + k1 + k2
