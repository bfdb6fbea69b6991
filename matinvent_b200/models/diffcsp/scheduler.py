"""Noise-schedule tables of the DiffCSP back-end (host side, built once per model).

Mirrors models/diffcsp/scheduler.py:7-116 of the reference: `BetaScheduler` / `SigmaScheduler` keep the
same buffer names (betas, alphas, alphas_cumprod, sigmas, sigmas_norm; index 0 is the t=0 pad) so a
reference checkpoint's `state_dict` loads unchanged.  The tables are fp32 torch tensors computed with
the reference's formulae; the per-step scalars the CUDA update kernels take are derived from them in
fp32 (`StepCoefficients`).  `sigmas_norm` is a Monte-Carlo estimate in the reference
(scheduler.py:46-51): load it from the checkpoint (or pass `sigmas_norm=`) for reproducible results.
"""
import math

import numpy as np
import torch
import torch.nn as nn


def _betas(timesteps, mode, beta_start, beta_end):
    if mode == "cosine":                                   # scheduler.py:7-16
        steps = torch.linspace(0, timesteps, timesteps + 1)
        f = torch.cos((steps / timesteps + 0.008) / 1.008 * math.pi * 0.5) ** 2
        f = f / f[0]
        return torch.clip(1 - f[1:] / f[:-1], 0.0001, 0.9999)
    if mode == "linear":
        return torch.linspace(beta_start, beta_end, timesteps)
    if mode == "quadratic":
        return torch.linspace(beta_start ** 0.5, beta_end ** 0.5, timesteps) ** 2
    if mode == "sigmoid":
        return torch.sigmoid(torch.linspace(-6, 6, timesteps)) * (beta_end - beta_start) + beta_start
    raise ValueError("unknown scheduler_mode %r" % (mode,))


class BetaScheduler(nn.Module):
    """DDPM variance schedule for lattices and atom types (scheduler.py:54-92)."""

    def __init__(self, timesteps, scheduler_mode, beta_start=0.0001, beta_end=0.02):
        super().__init__()
        self.timesteps = timesteps
        b = torch.cat([torch.zeros(1), _betas(timesteps, scheduler_mode, beta_start, beta_end)])
        a = 1.0 - b
        ac = torch.cumprod(a, 0)
        s = torch.zeros_like(b)
        s[1:] = b[1:] * (1.0 - ac[:-1]) / (1.0 - ac[1:])
        self.register_buffer("betas", b)
        self.register_buffer("alphas", a)
        self.register_buffer("alphas_cumprod", ac)
        self.register_buffer("sigmas", torch.sqrt(s))

    def uniform_sample_t(self, batch_size, device):
        return torch.from_numpy(np.random.choice(np.arange(1, self.timesteps + 1), batch_size)).to(device)


def wrapped_normal_score_sq_mean(sigmas, samples=10000, periods=10, generator=None):
    """E[(d log p_wn)^2] by Monte Carlo (scheduler.py:32-51), host side."""
    z = torch.randn(samples, sigmas.numel(), generator=generator)
    x = (sigmas * z) % 1.0
    s2 = sigmas ** 2
    num = torch.zeros_like(x)
    den = torch.zeros_like(x)
    for i in range(-periods, periods + 1):
        e = torch.exp(-(x + i) ** 2 / 2 / s2)
        num += (x + i) / s2 * e
        den += e
    return ((num / den) ** 2).mean(dim=0)


class SigmaScheduler(nn.Module):
    """Geometric sigma schedule of the wrapped-normal coordinate diffusion (scheduler.py:95-116)."""

    def __init__(self, timesteps, sigma_begin=0.01, sigma_end=1.0, sigmas_norm=None):
        super().__init__()
        self.timesteps, self.sigma_begin, self.sigma_end = timesteps, sigma_begin, sigma_end
        sig = torch.FloatTensor(np.exp(np.linspace(np.log(sigma_begin), np.log(sigma_end), timesteps)))
        if sigmas_norm is None:
            sn = torch.cat([torch.ones(1), wrapped_normal_score_sq_mean(sig)])
        else:
            sn = torch.as_tensor(sigmas_norm, dtype=torch.float32).detach().clone().cpu()
            if sn.numel() != timesteps + 1:
                raise ValueError("sigmas_norm must have timesteps+1 entries")
        self.register_buffer("sigmas", torch.cat([torch.zeros(1), sig]))
        self.register_buffer("sigmas_norm", sn)

    def uniform_sample_t(self, batch_size, device):
        return torch.from_numpy(np.random.choice(np.arange(1, self.timesteps + 1), batch_size)).to(device)


def time_frequencies(dim):
    """exp(-k ln(1e4)/(dim/2-1)), k < dim/2, as the reference builds it (diffusion.py:59-63)."""
    half = dim // 2
    e = math.log(10000) / (half - 1)
    return torch.exp(torch.arange(half) * -e)


def time_embedding_table(timesteps, dim):
    """Rows t = 0..T of SinusoidalTimeEmbeddings (diffusion.py:53-66), built on the host once."""
    t = torch.arange(timesteps + 1)
    arg = t[:, None] * time_frequencies(dim)[None, :]
    return torch.cat((arg.sin(), arg.cos()), dim=-1)


class StepCoefficients:
    """fp32 scalars of reverse step t (diffusion.py:300-307,324-325,341-343), same operation order."""

    def __init__(self, beta, sigma, step_lr):
        T = beta.timesteps
        al, ac = beta.alphas.cpu(), beta.alphas_cumprod.cpu()
        sx, sn = sigma.sigmas.cpu(), sigma.sigmas_norm.cpu()
        self.c0 = (1.0 / torch.sqrt(al)).tolist()
        self.c1 = ((1 - al) / torch.sqrt(1 - ac)).tolist()     # index 0 is 0/0 = nan, never used
        self.sig = beta.sigmas.cpu().tolist()
        self.sqrt_sn = torch.sqrt(sn).tolist()
        step_c = step_lr * (sx / sigma.sigma_begin) ** 2
        self.step_c = step_c.tolist()
        self.std_c = torch.sqrt(2 * step_c).tolist()
        adj = torch.cat([torch.zeros(1), sx[:-1]])
        self.step_p = (sx ** 2 - adj ** 2).tolist()
        std_p = torch.sqrt((adj ** 2 * (sx ** 2 - adj ** 2)) / (sx ** 2))
        self.std_p = std_p.tolist()
        self.T = T

    def table(self):
        """[T+1, 8] fp32 rows {sqrt_sn, step_c, std_c, step_p, std_p, c0, c1, sig} for the update kernels."""
        cols = [self.sqrt_sn, self.step_c, self.std_c, self.step_p, self.std_p, self.c0, self.c1, self.sig]
        return torch.tensor(cols, dtype=torch.float64).t().contiguous().to(torch.float32)
