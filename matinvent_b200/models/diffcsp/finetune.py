"""Fine-tune data + the fused reward-weighted update of the DiffCSP back-end.

`DiffCSPDataset` mirrors models/diffcsp/finetune.py:5-18 (attach one reward per crystal).
`FineTuner` is the device-side engine behind `MatInvent.ft_step` (pipeline/mat_invent.py:125-189): per
timestep  add_noise -> agent forward (activations kept) -> prior forward -> per-crystal loss + KL proxy +
reward weighting with analytic output gradients (mi_rl_loss) -> hand-written backward into ONE flat fp32
gradient buffer; every `accum_steps` timesteps one all-reduce of that buffer across ranks (NCCL over
NVLink; gloo in CPU tests of the host logic) followed by the flat Adam kernel.  A whole timestep is one
captured CUDA graph (the step index and the schedule scalars live on the device), replayed T times.

Sharding: every rank holds the full batch description, owns a contiguous slice of the
crystals balanced by sum n^2, draws the noise for the GLOBAL batch and slices it (so results do not
depend on the world size), scales its loss by 1/(B_global * accum_steps) and SUM-reduces gradients.
"""
import torch

from ... import ops
from .cspnet import MAX_ATOMIC_NUM
from .diffusion import PhiloxNoise, TorchNoise
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
    """Contiguous slices [start, end) per rank, balanced by sum n^2 (edge count)."""
    w = [int(n) * int(n) for n in num_atoms]
    total, B = sum(w), len(w)
    bounds, acc, r = [0], 0, 1
    for i, v in enumerate(w):
        acc += v
        while r < world and acc >= total * r / world and len(bounds) < world:
            bounds.append(i + 1)
            r += 1
    while len(bounds) < world:
        bounds.append(B)
    bounds.append(B)
    return [(bounds[k], max(bounds[k], bounds[k + 1])) for k in range(world)]


class FineTuner:
    def __init__(self, agent, prior, lr, accum_steps, sigma, process_group=None, rank=0, world=1, noise=None,
                 use_cuda_graph=True):
        self.agent, self.prior = agent, prior
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
        if hi <= lo:
            raise ValueError("rank %d got no crystals: need at least one crystal per rank" % self.rank)
        local = _LocalBatch(agent, batch, lo, hi, n_lo, n_hi)
        g = local.graph
        B, N = g.B, g.N
        reward = batch.reward.to(dev, torch.float32)[lo:hi].contiguous()
        w_kl = (self.sigma * (1.1 - reward)).contiguous()
        scale = 1.0 / (Bg * self.accum)
        noise = self.noise or TorchNoise(dev)
        in_graph_noise = isinstance(noise, PhiloxNoise)
        # global noise buffers (every rank draws the whole batch, uses its slice)
        z_l, z_x, z_a = (torch.empty(Bg, 3, 3, device=dev), torch.empty(Ng, 3, device=dev),
                         torch.empty(Ng, A, device=dev))
        zl, zx, za = z_l[lo:hi], z_x[n_lo:n_hi], z_a[n_lo:n_hi]
        l_t, x_t = torch.empty(B, 3, 3, device=dev), torch.empty(N, 3, device=dev)
        a_t, tar_x = torch.empty(N, A, device=dev), torch.empty(N, 3, device=dev)
        temb = torch.empty(B, agent.time_dim, device=dev)
        loss, kl = torch.empty(B, device=dev), torch.empty(B, device=dev)
        d = (torch.empty(B, 3, 3, device=dev), torch.empty(N, 3, device=dev), torch.empty(N, A, device=dev))
        stats = torch.zeros(2, device=dev)
        t_dev = torch.full((1,), T, dtype=torch.int32, device=dev)
        ttab, ntab = agent.time_table(), agent.noise_table()
        ws_a, ws_p = dec.workspace(g, True), pdec.workspace(g, False)
        costs = agent._costs()

        def body():
            ops.sampler_step_begin(t_dev, ttab, temb, B, agent.time_dim)
            if in_graph_noise:
                noise.fill(z_l), noise.fill(z_x), noise.fill(z_a)      # draw order :102, :111
            ops.add_noise(local.L0, local.x0, local.Z, zl, zx, za, B, N, A, ntab, l_t, x_t, a_t, tar_x, t_dev=t_dev)
            pa = dec.forward_graph(g, temb, a_t, x_t, l_t, train=True, ws=ws_a)
            pp = pdec.forward_graph(g, temb, a_t, x_t, l_t, train=False, ws=ws_p)
            ops.rl_loss(pa, (zl, tar_x, za), pp, g.node_off, B, A, costs, reward, w_kl, scale, loss, kl, d, stats)
            dec.backward_graph(g, temb, a_t, x_t, l_t, d[0], d[1], d[2], ws=ws_a)
            ops.sampler_step_end(t_dev)

        graph = None
        for t in range(timesteps):
            if not in_graph_noise:
                z_l.copy_(noise.step_randn((Bg, 3, 3)))
                z_x.copy_(noise.step_randn((Ng, 3)))
                z_a.copy_(noise.step_randn((Ng, A)))
            if not self.use_graph or t == 0:
                body()
            else:
                if graph is None:
                    torch.cuda.synchronize()
                    graph = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(graph):
                        body()
                graph.replay()
            if (t + 1) % self.accum == 0:
                self.optimizer_step()
        if timesteps % self.accum != 0:
            self.optimizer_step()
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
