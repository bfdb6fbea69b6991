"""Fine-tune data + the fused reward-weighted update of the DiffCSP back-end.

`DiffCSPDataset` mirrors models/diffcsp/finetune.py:5-18 (attach one reward per crystal).
`FineTuner` is the device-side engine behind `MatInvent.ft_step` (pipeline/mat_invent.py:125-189): per
timestep  add_noise -> agent forward (activations kept) -> prior forward -> per-crystal loss + KL proxy +
reward weighting with analytic output gradients (mi_rl_loss) -> hand-written backward into ONE flat fp32
gradient buffer; every `accum_steps` timesteps one all-reduce of that buffer across ranks (NCCL over
NVLink; gloo in CPU tests of the host logic) followed by the flat Adam kernel.

Timestep grouping.  The reference fine-tunes <= 18 crystals for 3 x 1000 timesteps: one timestep of
such a batch is ~260 launches of a few microseconds of work each (5.2 ms, launch- and latency-bound).  The weights
only change every `accum_steps` timesteps, so the timesteps inside one accumulation window are independent: G of
them are stacked into ONE batch of G x B crystals (each copy with its own time index and noise) and go through one
forward / backward.  The gradient is the same sum, accumulated in a different order (fp32: ~1e-7 relative); the
noise is drawn slot by slot in the reference's order.  A whole group is one captured CUDA graph (the time indices
and schedule scalars live on the device), replayed timesteps / G times.

Sharding: every rank holds the full batch description, owns a contiguous slice of the
crystals balanced by sum n^2, draws the noise for the GLOBAL batch and slices it (so results do not
depend on the world size), scales its loss by 1/(B_global * accum_steps) and SUM-reduces gradients.
"""
import torch

from ... import ops
from .cspnet import MAX_ATOMIC_NUM
from .diffusion import PhiloxNoise, TorchNoise, capture_graph
from .sample import CrystalBatch


class DiffCSPDataset:
    def __init__(self, data_list, rewards=None):
        self.data_list = data_list
        if rewards is not None:
            rewards = torch.as_tensor(rewards, dtype=torch.float)
            for i, data in enumerate(self.data_list):
                data.reward = rewards[i].unsqueeze(dim=0)

    def __len__(self):
        return len(self.data_list)

    def __getitem__(self, index):
        return self.data_list[index]


class CrystalLoader:
    """Stand-in for the PyG DataLoader of models/suite/diffcsp.py:116-131 (shuffle + collate)."""

    def __init__(self, dataset, batch_size, shuffle=True):
        self.dataset, self.batch_size, self.shuffle = dataset, batch_size, shuffle

    def __len__(self):
        return (len(self.dataset) + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        n = len(self.dataset)
        order = torch.randperm(n).tolist() if self.shuffle else list(range(n))
        for i in range(0, n, self.batch_size):
            yield CrystalBatch([self.dataset[j] for j in order[i:i + self.batch_size]])


def partition_crystals(num_atoms, world):
    """Contiguous slices [start, end) per rank, balanced by sum n^2 (edge count).  Every rank gets at least one crystal
    when there are at least `world` of them (a 20-atom crystal next to 1-atom ones may cross several of the
    total * r / world thresholds at once); with fewer crystals than ranks the trailing ranks get empty slices — they
    contribute a zero gradient and still join every collective (FineTuner.run_batch)."""
    w = [int(n) * int(n) for n in num_atoms]
    B, total = len(w), sum(w)
    if B < world:                              # one crystal each for the first B ranks
        return [(min(r, B), min(r + 1, B)) for r in range(world)]
    prefix = [0]
    for v in w:
        prefix.append(prefix[-1] + v)
    bounds = [0]
    b = 0
    for r in range(1, world):
        target = total * r / world
        while b < B and prefix[b] < target:
            b += 1
        if b > 0 and target - prefix[b - 1] < prefix[b] - target:        # the nearer of the two cuts
            b -= 1
        b = min(max(b, bounds[-1] + 1), B - (world - r))                  # strictly increasing, one left for each later rank
        bounds.append(b)
    bounds.append(B)
    return [(bounds[k], bounds[k + 1]) for k in range(world)]


class FineTuner:
    def __init__(self, agent, prior, lr, accum_steps, sigma, process_group=None, rank=0, world=1, noise=None,
                 use_cuda_graph=True, group=None):
        self.agent, self.prior = agent, prior
        self.group = group          # timesteps stacked per launch; None = chosen from the batch's edge count
        self.lr, self.accum, self.sigma = float(lr), int(accum_steps), float(sigma)
        self.pg, self.rank, self.world = process_group, rank, world
        self.noise = noise
        self.use_graph = use_cuda_graph and agent.decoder.edge_style == "fc"
        dec = agent.decoder
        self.grad = dec.flat_grad()
        self.grad.zero_()
        # fresh Adam state per ft_step, like torch.optim.Adam(...) at pipeline/mat_invent.py:136
        self.m = torch.zeros_like(self.grad)
        self.v = torch.zeros_like(self.grad)
        self.adam_t = 0

    # ------------------------------------------------------------------ one epoch over one batch
    # edges per stacked launch the grouping aims for.  Measured at the reference's 18 crystals (2 786 edges), ms per
    # timestep: G=1 5.2, G=5 1.9, G=10 0.78, G=25 0.66, G=50 0.60 — the weight-gradient GEMMs go to the tensor cores from
    # ~20k reduction rows (CSPNet.WGRAD_TC_ROWS) and like long reductions; the training workspace is ~70 KB per edge
    GROUP_EDGES = 150000

    def group_size(self, edges_local):
        """largest divisor of accum_steps that keeps a stacked group at <= GROUP_EDGES edges (>= 1)"""
        want = max(1, min(self.accum, self.GROUP_EDGES // max(1, edges_local)))
        if self.group is not None:
            want = max(1, min(self.accum, int(self.group)))
        return max(d for d in range(1, want + 1) if self.accum % d == 0)

    def run_batch(self, batch, timesteps):
        """for t in range(timesteps): ... (pipeline/mat_invent.py:150-170).  Returns the epoch's
        (loss, loss_diff, loss_kl) sums in the reference's normalisation (per-batch, :172-174)."""
        agent, prior = self.agent, self.prior
        dev = agent.device
        dec, pdec = agent.decoder, prior.decoder
        A = MAX_ATOMIC_NUM
        T = agent.timesteps
        na = batch.num_atoms.tolist()
        Bg, Ng = len(na), sum(na)
        lo, hi = partition_crystals(na, self.world)[self.rank]
        n_lo, n_hi = sum(na[:lo]), sum(na[:hi])
        empty = hi <= lo        # fewer crystals than ranks: this rank adds a zero gradient and joins every collective
        local = None if empty else _LocalBatch(agent, batch, lo, hi, n_lo, n_hi)
        B, N = (0, 0) if empty else (local.graph.B, local.graph.N)
        reward = batch.reward.to(dev, torch.float32)[lo:hi].contiguous()
        w_kl = (self.sigma * (1.1 - reward)).contiguous()
        scale = 1.0 / (Bg * self.accum)
        noise = self.noise or TorchNoise(dev)
        in_graph_noise = isinstance(noise, PhiloxNoise)
        ttab, ntab = agent.time_table(), agent.noise_table()
        costs = agent._costs()
        stats = torch.zeros(2, device=dev)
        # every rank must cut the epoch into the same groups (noise is drawn per group for the GLOBAL batch): size them
        # from the largest shard
        parts = partition_crystals(na, self.world)
        e_max = max(sum(n * n for n in na[a:b]) for a, b in parts)
        G = self.group_size(e_max) if self.use_graph or self.group else 1
        groups = {}
        used_graphs = []

        def make_group(Gn):
            """buffers + body for Gn stacked timesteps (slot j holds timestep t0 + j)"""
            z_l, z_x, z_a = (torch.empty(Gn, Bg, 3, 3, device=dev), torch.empty(Gn, Ng, 3, device=dev),
                             torch.empty(Gn, Ng, A, device=dev))
            if empty:
                def draw_only():
                    if in_graph_noise:
                        for j in range(Gn):
                            noise.fill(z_l[j]), noise.fill(z_x[j]), noise.fill(z_a[j])
                return dict(body=draw_only, z=(z_l, z_x, z_a), t_vec=torch.zeros(Gn, dtype=torch.int32, device=dev),
                            graph=None, runs=0, eager=True)
            g = dec.graph_for(na[lo:hi] * Gn)
            used_graphs.append(g)
            # global noise per slot (every rank draws the whole batch, uses its slice)
            if self.world > 1:      # this rank's crystals of every slot, contiguous (loss targets)
                zl, zx, za = (torch.empty(Gn, B, 3, 3, device=dev), torch.empty(Gn, N, 3, device=dev),
                              torch.empty(Gn, N, A, device=dev))
            else:
                zl, zx, za = z_l, z_x, z_a
            l_t, x_t = torch.empty(Gn * B, 3, 3, device=dev), torch.empty(Gn * N, 3, device=dev)
            a_t, tar_x = torch.empty(Gn * N, A, device=dev), torch.empty(Gn * N, 3, device=dev)
            temb = torch.empty(Gn * B, agent.time_dim, device=dev)
            loss, kl = torch.empty(Gn * B, device=dev), torch.empty(Gn * B, device=dev)
            d = (torch.empty(Gn * B, 3, 3, device=dev), torch.empty(Gn * N, 3, device=dev), torch.empty(Gn * N, A, device=dev))
            t_vec = torch.zeros(Gn, dtype=torch.int32, device=dev)          # time index of every slot
            rew, wk = reward.repeat(Gn).contiguous(), w_kl.repeat(Gn).contiguous()
            ws_a, ws_p = dec.workspace(g, True), pdec.workspace(g, False)

            def body():
                for j in range(Gn):
                    tj = t_vec[j:j + 1]
                    ops.sampler_step_begin(tj, ttab, temb[j * B:(j + 1) * B], B, agent.time_dim)
                    if in_graph_noise:
                        noise.fill(z_l[j]), noise.fill(z_x[j]), noise.fill(z_a[j])      # draw order :102, :111
                    if self.world > 1:
                        zl[j].copy_(z_l[j, lo:hi]), zx[j].copy_(z_x[j, n_lo:n_hi]), za[j].copy_(z_a[j, n_lo:n_hi])
                    ops.add_noise(local.L0, local.x0, local.Z, zl[j], zx[j], za[j], B, N, A, ntab,
                                  l_t[j * B:(j + 1) * B], x_t[j * N:(j + 1) * N], a_t[j * N:(j + 1) * N],
                                  tar_x[j * N:(j + 1) * N], t_dev=tj)
                pa = dec.forward_graph(g, temb, a_t, x_t, l_t, train=True, ws=ws_a)
                pp = pdec.forward_graph(g, temb, a_t, x_t, l_t, train=False, ws=ws_p)
                ops.rl_loss(pa, (zl.view(Gn * B, 3, 3), tar_x, za.view(Gn * N, A)), pp, g.node_off, Gn * B, A, costs, rew, wk,
                            scale, loss, kl, d, stats)
                dec.backward_graph(g, temb, a_t, x_t, l_t, d[0], d[1], d[2], ws=ws_a)

            return dict(body=body, z=(z_l, z_x, z_a), t_vec=t_vec, graph=None, runs=0, eager=False)

        t = 0
        while t < timesteps:
            Gn = min(G, self.accum - (t % self.accum), timesteps - t)     # a group never straddles an optimizer step
            grp = groups.get(Gn)
            if grp is None:
                grp = groups[Gn] = make_group(Gn)
            z_l, z_x, z_a = grp["z"]
            # time of timestep it: T - it (times = T - t_idx, diffusion.py:86-87)
            grp["t_vec"].copy_(torch.arange(T - t, T - t - Gn, -1, dtype=torch.int32), non_blocking=False)
            if not in_graph_noise:
                for j in range(Gn):
                    z_l[j].copy_(noise.step_randn((Bg, 3, 3)))
                    z_x[j].copy_(noise.step_randn((Ng, 3)))
                    z_a[j].copy_(noise.step_randn((Ng, A)))
            if not self.use_graph or grp["runs"] == 0 or grp["eager"]:
                grp["body"]()
            else:
                if grp["graph"] is None:
                    grp["graph"] = capture_graph(grp["body"])
                grp["graph"].replay()
            grp["runs"] += 1
            t += Gn
            if t % self.accum == 0:
                self.optimizer_step()
        if timesteps % self.accum != 0:
            self.optimizer_step()
        # the stacked graphs and their workspaces (~70 KB per edge for the agent) belong to this batch order only: the
        # loader reshuffles every epoch, so they are released here instead of piling up in the decoders' caches
        groups.clear()
        for g_ in used_graphs:
            dec.release(g_), pdec.release(g_)
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(stats, group=self.pg)
        s = stats.tolist()
        loss_diff = s[0] / timesteps
        loss_kl = s[1] / self.sigma / timesteps if self.sigma != 0 else 0.0
        loss_mean = (s[0] + s[1]) / Bg / timesteps
        return loss_mean, loss_diff, loss_kl

    def optimizer_step(self):
        """optimizer.step(); optimizer.zero_grad() (pipeline/mat_invent.py:165-167) with the gradient summed
        over ranks first."""
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(self.grad, group=self.pg)
        self.adam_t += 1
        ops.adam_step(self.agent.decoder.flat.data, self.grad, self.m, self.v, self.lr, self.adam_t, zero_grad=True)
        self.agent.decoder.weights_changed()


class _LocalBatch:
    """This rank's contiguous slice of a collated batch, on the device."""

    def __init__(self, module, batch, lo, hi, n_lo, n_hi):
        dev = module.device
        self.graph = module.decoder.graph_for(batch.num_atoms[lo:hi])
        g = self.graph
        f = lambda t: t.to(dev, torch.float32).contiguous()
        self.x0 = f(batch.frac_coords[n_lo:n_hi]).view(g.N, 3)
        self.Z = batch.atom_types[n_lo:n_hi].to(dev, torch.int32).contiguous().view(g.N)
        lengths, angles = f(batch.lengths[lo:hi]).view(g.B, 3), f(batch.angles[lo:hi]).view(g.B, 3)
        self.L0 = torch.empty(g.B, 3, 3, device=dev)
        ops.lattice_params_to_matrix(lengths, angles, self.L0, g.B)
