"""DiffCSP diffusion module on libmatinvent_b200: forward noising, per-crystal losses, KL proxy and the
1000-step predictor-corrector reverse sampler.

Host-side mirror of the reference's `DiffCSPModule` (models/diffcsp/diffusion.py:69-399) with the same
method names and return structure — `add_noise(batch, t)`, `calc_sample_loss(noised)`,
`calc_kl_reg(agent_pred, prior_pred, batch)`, `sample(batch, step_lr=...)` — so it can be driven by
the unmodified `MatInvent.ft_step` (pipeline/mat_invent.py:125-189) and `DiffCSPSampler.generate`.
Differences (cost only, not results): the reverse loop launches a captured CUDA graph per step, keeps
only the current state instead of 1001 trajectory entries (diffusion.py:384-390), skips the four
log-probabilities nobody reads (:353-382) and the lattice/type heads of the corrector forward whose
outputs the reference discards (:327).  Noise is drawn in the reference's order (:277-279, 320-322,
337-339) from a pluggable source so CPU-oracle and GPU runs can share a tape.
"""
import numpy as np
import torch
import torch.nn as nn

from ... import ops
from .cspnet import CSPNet, MAX_ATOMIC_NUM
from .scheduler import BetaScheduler, SigmaScheduler, StepCoefficients, time_embedding_table


class TorchNoise:
    """Default noise source: torch's global generators, exactly like the reference — the initial state
    on the CPU generator (diffusion.py:277-279), per-step noise with randn on the model's device
    (:320-322, 337-339; two of every three tensors are unused but still advance the generator)."""

    def __init__(self, device):
        self.device = device

    def init_rand(self, shape):
        return torch.rand(shape).to(self.device)

    def init_randn(self, shape):
        return torch.randn(shape).to(self.device)

    def step_randn(self, shape, used=True):
        return torch.randn(shape, device=self.device)


class TapeNoise:
    """Noise from a CPU torch.Generator in the reference's draw order (parity tests share it with the
    oracle).  Unused draws still consume the generator."""

    def __init__(self, device, seed=None, generator=None):
        self.device = device
        self.g = generator if generator is not None else torch.Generator().manual_seed(seed)

    def init_rand(self, shape):
        return torch.rand(tuple(shape), generator=self.g).to(self.device)

    def init_randn(self, shape):
        return torch.randn(tuple(shape), generator=self.g).to(self.device)

    def step_randn(self, shape, used=True):
        t = torch.randn(tuple(shape), generator=self.g)
        return t.to(self.device) if used else None


class PhiloxNoise:
    """In-library Philox4x32-10 normals written straight into the step buffers (graph-capturable):
    the throughput path.  Not stream-compatible with torch's generators; unused draws are skipped."""

    def __init__(self, device, seed=0):
        self.device, self.seed = device, int(seed)
        self.offset = torch.zeros(1, dtype=torch.int64, device=device)

    def init_rand(self, shape):
        out = torch.empty(tuple(shape), device=self.device)
        return ops.philox_uniform(out, self.seed, 0, self.offset, True)

    def init_randn(self, shape):
        out = torch.empty(tuple(shape), device=self.device)
        return ops.philox_normal(out, self.seed, 0, self.offset, True)

    def fill(self, out):
        return ops.philox_normal(out, self.seed, 0, self.offset, True)

    def step_randn(self, shape, used=True):
        if not used:
            return None
        return self.init_randn(shape)


class _Hparams(dict):
    __getattr__ = dict.__getitem__

    def __setattr__(self, k, v):
        self[k] = v


class _Batch:
    """Device-side view of a PyG-style batch (num_atoms, batch, frac_coords, atom_types, lengths, angles)."""

    def __init__(self, module, batch):
        dev = module.device
        self.graph = module.decoder.graph_for(batch.num_atoms)
        g = self.graph
        f = lambda t: t.to(dev, torch.float32).contiguous()
        self.x0 = f(batch.frac_coords).view(g.N, 3)
        self.Z = batch.atom_types.to(dev, torch.int32).contiguous().view(g.N)
        self.lengths, self.angles = f(batch.lengths).view(g.B, 3), f(batch.angles).view(g.B, 3)
        self.L0 = torch.empty(g.B, 3, 3, device=dev)
        ops.lattice_params_to_matrix(self.lengths, self.angles, self.L0, g.B)


class DiffCSPModule(nn.Module):
    """`decoder` = CSPNet(**decoder_cfg, latent_dim=latent_dim+time_dim, pred_type=True, smooth=True)
    (diffusion.py:73).  Constructor takes plain dicts instead of hydra nodes:

        DiffCSPModule(decoder=dict(hidden_dim=512, num_layers=6, ...), beta_scheduler=dict(timesteps=1000,
                      scheduler_mode='cosine'), sigma_scheduler=dict(timesteps=1000, sigma_begin=0.005,
                      sigma_end=0.5), cost_lattice=1., cost_coord=1., cost_type=20., time_dim=256, latent_dim=0)
    """

    def __init__(self, decoder, beta_scheduler, sigma_scheduler, cost_lattice=1.0, cost_coord=1.0, cost_type=20.0,
                 time_dim=256, latent_dim=0, device=None, sigmas_norm=None, **kwargs):
        super().__init__()
        dev = torch.device(device if device is not None else "cuda")
        self.hparams = _Hparams(cost_lattice=cost_lattice, cost_coord=cost_coord, cost_type=cost_type,
                                time_dim=time_dim, latent_dim=latent_dim, decoder=dict(decoder),
                                beta_scheduler=dict(beta_scheduler), sigma_scheduler=dict(sigma_scheduler), **kwargs)
        dec = {k: v for k, v in dict(decoder).items() if k not in ("_target_", "latent_dim", "pred_type", "smooth")}
        self.decoder = CSPNet(**dec, latent_dim=latent_dim + time_dim, pred_type=True, smooth=True, device=dev)
        bs = {k: v for k, v in dict(beta_scheduler).items() if k != "_target_"}
        ss = {k: v for k, v in dict(sigma_scheduler).items() if k != "_target_"}
        self.beta_scheduler = BetaScheduler(**bs)
        self.sigma_scheduler = SigmaScheduler(**ss, sigmas_norm=sigmas_norm)
        self.time_dim = time_dim
        self.keep_lattice = cost_lattice < 1e-5
        self.keep_coords = cost_coord < 1e-5
        if self.keep_lattice or self.keep_coords:
            raise NotImplementedError("keep_lattice / keep_coords (zero lattice/coord cost) is not part of the MatInvent path")
        self._time_table = None
        self._noise_table = None
        self._step_graphs = {}
        self.to(dev)

    # ------------------------------------------------------------------ plumbing
    @property
    def device(self):
        return self.decoder.device

    @property
    def timesteps(self):
        return self.beta_scheduler.timesteps

    def time_table(self):
        """[T+1, time_dim] SinusoidalTimeEmbeddings rows (diffusion.py:53-66), device resident."""
        if self._time_table is None or self._time_table.device != self.device:
            self._time_table = time_embedding_table(self.timesteps, self.time_dim).to(self.device)
        return self._time_table

    def noise_table(self):
        """[T+1, 4] rows {sqrt(abar_t), sqrt(1-abar_t), sigma_t, sqrt(sigma_norm_t)} of add_noise (diffusion.py:91-98)."""
        if self._noise_table is None or self._noise_table.device != self.device:
            ac = self.beta_scheduler.alphas_cumprod.cpu()
            cols = [torch.sqrt(ac), torch.sqrt(1.0 - ac), self.sigma_scheduler.sigmas.cpu(),
                    torch.sqrt(self.sigma_scheduler.sigmas_norm.cpu())]
            self._noise_table = torch.stack(cols, dim=1).contiguous().to(self.device)
        return self._noise_table

    def _apply(self, fn, *a, **k):
        r = super()._apply(fn, *a, **k)
        self._time_table = None
        self._noise_table = None
        self._step_graphs = {}
        return r

    def _costs(self):
        hp = self.hparams
        return float(hp.cost_lattice), float(hp.cost_coord), float(hp.cost_type)

    # ------------------------------------------------------------------ forward noising (diffusion.py:81-119)
    def add_noise(self, batch, time=None, noise=None):
        if time is None:
            raise NotImplementedError("random per-crystal times are a training-from-scratch feature; MatInvent.ft_step passes t")
        T = self.timesteps
        t = int(np.arange(T, 0, -1)[time])
        dev = self.device
        db = batch if isinstance(batch, _Batch) else _Batch(self, batch)
        g = db.graph
        B, N, A = g.B, g.N, MAX_ATOMIC_NUM
        noise = noise or TorchNoise(dev)
        # draw order of the reference: rand_l, rand_x (:102) then rand_t (:111)
        z_l = noise.step_randn((B, 3, 3))
        z_x = noise.step_randn((N, 3))
        z_a = noise.step_randn((N, A))
        l_t, x_t = torch.empty(B, 3, 3, device=dev), torch.empty(N, 3, device=dev)
        a_t, tar_x = torch.empty(N, A, device=dev), torch.empty(N, 3, device=dev)
        ops.add_noise(db.L0, db.x0, db.Z, z_l, z_x, z_a, B, N, A, self.noise_table(), l_t, x_t, a_t, tar_x, t_host=t)
        temb = self.time_table()[t].expand(B, -1).contiguous()
        noised_input = (temb, a_t, x_t, l_t, g.num_atoms, g.node2graph)
        return noised_input, (z_l, tar_x, z_a), g.node2graph

    # ------------------------------------------------------------------ losses (diffusion.py:121-149)
    def calc_sample_loss(self, input_all):
        noised_input, (rand_l, tar_x, rand_t), _ = input_all
        pred = self.decoder(*noised_input)
        g = self.decoder.graph_for(noised_input[4])
        loss = _SampleLoss.apply(self, g, pred[0], pred[1], pred[2], rand_l, tar_x, rand_t, None, None, None)
        return loss, pred

    def calc_kl_reg(self, agent_pred, prior_pred, batch):
        g = self.decoder.graph_for(batch.num_atoms)
        pl, px, pa = agent_pred
        ql, qx, qa = (t.detach() for t in prior_pred)
        return _SampleLoss.apply(self, g, pl, px, pa, None, None, None, ql, qx, qa)

    # ------------------------------------------------------------------ reverse sampler (diffusion.py:273-399)
    @torch.no_grad()
    def sample(self, batch, diff_ratio=1.0, step_lr=1e-5, noise=None, use_cuda_graph=True, timesteps=None,
               return_traj=False):
        """Returns (traj[0], traj) like the reference; `traj` only holds the entries that were asked for
        (`return_traj=True` keeps every step, otherwise just the first and 0).  `timesteps` < T starts the
        loop at t = timesteps (bounded runs for tests / CPU-baseline comparison)."""
        dev = self.device
        g = self.decoder.graph_for(batch.num_atoms)
        B, N, A = g.B, g.N, MAX_ATOMIC_NUM
        noise = noise or TorchNoise(dev)
        T0 = self.timesteps if timesteps is None else int(timesteps)
        # initial state, reference draw order x_T, l_T, t_T (:277-279)
        x = noise.init_rand((N, 3)).to(torch.float32).contiguous()
        l = noise.init_randn((B, 3, 3)).to(torch.float32).contiguous()
        a = noise.init_randn((N, A)).to(torch.float32).contiguous()
        x.remainder_(1.0)

        def snap(st=None):
            src = st or _Snap(x, l, a)
            return dict(atom_types=src.a.clone(), frac_coords=src.x.clone(), lattices=src.l.clone(),
                        num_atoms=g.num_atoms, batch_idx=g.node2graph)

        traj = {T0: snap()}
        co = StepCoefficients(self.beta_scheduler, self.sigma_scheduler, step_lr)
        st = _StepState(self, g, x, l, a, co, T0, noise, use_cuda_graph and self.decoder.edge_style == "fc")
        for t in range(T0, 0, -1):
            st.step(last=(t == 1))
            if return_traj and t > 1:
                traj[t - 1] = snap(st)
        traj[0] = snap(st)
        return traj[0], traj


class _Snap:
    def __init__(self, x, l, a):
        self.x, self.l, self.a = x, l, a


class _StepState:
    """State + static buffers of a sampling run.  One reverse step = corrector forward + update, predictor
    forward + update (2 score-network evaluations).  The step index lives on the device (`t_dev`) and the
    per-step scalars in a device table, so ONE captured CUDA graph is replayed for every step; with
    PhiloxNoise the noise draws are part of the graph too (no host work between replays)."""

    def __init__(self, module, g, x, l, a, co, T0, noise, use_graph):
        dev = module.device
        self.m, self.g, self.noise = module, g, noise
        B, N, A = g.B, g.N, MAX_ATOMIC_NUM
        self.x, self.l, self.a = x, l, a
        self.x_half = torch.empty_like(x)
        self.temb = torch.empty(B, module.time_dim, device=dev)
        # the four noise draws of a step are views of ONE buffer: with in-graph Philox noise a step costs one fill
        sizes = [N * 3, B * 9, N * A, N * 3]
        offs = [0]
        for n_ in sizes:
            offs.append(offs[-1] + (n_ + 3) // 4 * 4)
        self.znoise = torch.zeros(offs[-1], device=dev)
        self.zx_c, self.zl = self.znoise[offs[0]:offs[0] + N * 3].view(N, 3), self.znoise[offs[1]:offs[1] + B * 9].view(B, 3, 3)
        self.za, self.zx_p = self.znoise[offs[2]:offs[2] + N * A].view(N, A), self.znoise[offs[3]:offs[3] + N * 3].view(N, 3)
        self.ws = module.decoder.workspace(g, False)
        self.coef = co.table().to(dev)
        self.ttab = module.time_table()
        # all crystals of a reverse step share the time: the per-crystal time term of the embedding for every timestep, once
        self.tbtab = module.decoder.time_term_table(self.ttab, a)
        self.t_dev = torch.full((1,), T0, dtype=torch.int32, device=dev)
        self.in_graph_noise = isinstance(noise, PhiloxNoise)
        self.use_graph = use_graph
        self.graph = None
        self.steps_done = 0

    def _host_noise(self):
        """Reference draw order per step (:320-322, 337-339): l, t, x (corrector; l and t unused), l, t, x."""
        g, nz = self.g, self.noise
        B, N, A = g.B, g.N, MAX_ATOMIC_NUM
        nz.step_randn((B, 3, 3), used=False)
        nz.step_randn((N, A), used=False)
        self.zx_c.copy_(nz.step_randn((N, 3)))
        self.zl.copy_(nz.step_randn((B, 3, 3)))
        self.za.copy_(nz.step_randn((N, A)))
        self.zx_p.copy_(nz.step_randn((N, 3)))

    def _body(self, with_noise):
        m, g, ws = self.m, self.g, self.ws
        B, N, A = g.B, g.N, MAX_ATOMIC_NUM
        dec = m.decoder
        ops.sampler_step_begin(self.t_dev, self.tbtab, ws.tb, B, dec.hidden_dim)     # ws.tb = row t of the time-term table
        if with_noise and self.in_graph_noise:
            self.noise.fill(self.znoise)
        nz = (lambda t: t) if with_noise else (lambda t: None)
        # corrector: only the coordinate head is consumed (diffusion.py:327-330)
        _, px, _ = dec.forward_graph(g, self.temb, self.a, self.x, self.l, heads=(False, True, False), ws=ws, tb_ready=True)
        ops.reverse_corrector(self.x, px, nz(self.zx_c), self.x_half, N, self.coef, self.t_dev)
        # predictor (diffusion.py:345-351)
        pl, px, pa = dec.forward_graph(g, self.temb, self.a, self.x_half, self.l, ws=ws, reuse_embedding=True, tb_ready=True)
        ops.reverse_predictor(self.x_half, px, nz(self.zx_p), self.x, N, self.l, pl, nz(self.zl), B, self.a, pa,
                              nz(self.za), A, self.coef, self.t_dev)
        ops.sampler_step_end(self.t_dev)

    def step(self, last):
        if not last and not self.in_graph_noise:
            self._host_noise()
        if last or not self.use_graph or self.steps_done == 0:
            self._body(with_noise=not last)          # first step doubles as warm-up for the capture
        else:
            if self.graph is None:
                self.graph = capture_graph(lambda: self._body(with_noise=True))
            self.graph.replay()
        self.steps_done += 1


def capture_graph(body):
    """One CUDA graph of `body()`, captured on a side stream with capture_begin / capture_end by hand: the
    `torch.cuda.graph` context also runs gc.collect(), torch.cuda.empty_cache() and a device synchronize — 0.1-0.15 s per
    sampled batch / fine-tune epoch (scripts/probe_e2e.py), 4-5 % of a 256-crystal sampling pass."""
    graph = torch.cuda.CUDAGraph()
    cur = torch.cuda.current_stream()
    side = torch.cuda.Stream()
    side.wait_stream(cur)
    with torch.cuda.stream(side):
        graph.capture_begin()
        try:
            body()
        finally:
            graph.capture_end()
    cur.wait_stream(side)
    return graph


class _SampleLoss(torch.autograd.Function):
    """Per-crystal MSE losses (diffusion.py:121-138) or the KL proxy (:140-149) via mi_rl_loss, with the
    analytic gradient w.r.t. the agent's predictions."""

    @staticmethod
    def forward(ctx, module, g, pl, px, pa, tl, tx, ta, ql, qx, qa):
        dev = module.device
        B = g.B
        c = lambda t: None if t is None else t.detach().to(torch.float32).contiguous()
        pred = (c(pl), c(px), c(pa))
        tgt = None if tl is None else (c(tl), c(tx), c(ta))
        prior = None if ql is None else (c(ql), c(qx), c(qa))
        loss, kl = torch.empty(B, device=dev), torch.empty(B, device=dev)
        ops.rl_loss(pred, tgt, prior, g.node_off, B, MAX_ATOMIC_NUM, module._costs(), None, None, 1.0, loss, kl, None)
        ctx.module, ctx.g, ctx.pred, ctx.tgt, ctx.prior = module, g, pred, tgt, prior
        return loss if tgt is not None else kl

    @staticmethod
    def backward(ctx, gout):
        module, g = ctx.module, ctx.g
        pl, px, pa = ctx.pred
        d = (torch.empty_like(pl), torch.empty_like(px), torch.empty_like(pa))
        gout = gout.contiguous()
        if ctx.tgt is not None:
            ops.rl_loss(ctx.pred, ctx.tgt, None, g.node_off, g.B, MAX_ATOMIC_NUM, module._costs(), gout, None, 1.0,
                        None, None, d)
        else:
            ops.rl_loss(ctx.pred, None, ctx.prior, g.node_off, g.B, MAX_ATOMIC_NUM, module._costs(), None, gout, 1.0,
                        None, None, d)
        return (None, None, d[0], d[1], d[2], None, None, None, None, None, None)
