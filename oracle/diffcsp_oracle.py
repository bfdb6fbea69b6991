"""ORACLE — TEST INFRASTRUCTURE ONLY.  CPU restatement (torch CPU ops, fp32 or fp64) of the reference's
DiffCSP hot path: score network, noise schedules, forward noising, per-crystal losses, the 1000-step
predictor-corrector reverse sampler, the reward-weighted fine-tune step and `generate` post-processing.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may
import this file; the product path (`matinvent_b200/`) never does and fails loudly without its CUDA
library.

PINNING: every function below is checked against the UNMODIFIED reference files imported under
`oracle/shims` (tests/test_oracle_vs_reference.py, runs wherever /root/reference exists) and against
the golden vectors those reference files produced (tests/golden/*.pt, made by oracle/make_golden.py).
The reference ships no tests/golden vectors of its own except the `repeat_blocks` docstring examples
(models/diffcsp/utils.py:208-226), which tests/test_oracle_golden.py holds as known answers.

All `file:line` citations are relative to /root/reference.
"""
import math

import numpy as np
import torch

MAX_ATOMIC_NUM = 100

# models/diffcsp/sample.py:42-62 — atom-count prior for mp_20 (index = number of atoms).
ATOM_DIST_MP20 = [0.0, 0.0021742334905660377, 0.021079009433962265, 0.019826061320754717,
                  0.15271226415094338, 0.047132959905660375, 0.08464770047169812, 0.021079009433962265,
                  0.07808814858490566, 0.03434551886792453, 0.0972877358490566, 0.013303360849056603,
                  0.09669811320754718, 0.02155807783018868, 0.06522700471698113, 0.014372051886792452,
                  0.06703272405660378, 0.00972877358490566, 0.053176591981132074, 0.010576356132075472,
                  0.08995430424528301]


def default_hparams(**over):
    """Upstream-DiffCSP default sizes; the real values ride in the checkpoint's
    hparams.yaml (models/suite/diffcsp.py:52-63)."""
    hp = dict(hidden_dim=512, latent_dim=0, time_dim=256, num_layers=6, max_atoms=100, num_freqs=128,
              edge_style="fc", cutoff=7.0, max_neighbors=20, ln=True, ip=True,
              timesteps=1000, beta_mode="cosine", sigma_begin=0.005, sigma_end=0.5,
              cost_lattice=1.0, cost_coord=1.0, cost_type=20.0)
    hp.update(over)
    return hp


# ----------------------------------------------------------------------------- schedules
def cosine_beta_schedule(timesteps, s=0.008):
    """models/diffcsp/scheduler.py:7-16."""
    steps = timesteps + 1
    x = torch.linspace(0, timesteps, steps)
    ac = torch.cos(((x / timesteps) + s) / (1 + s) * math.pi * 0.5) ** 2
    ac = ac / ac[0]
    betas = 1 - (ac[1:] / ac[:-1])
    return torch.clip(betas, 0.0001, 0.9999)


def beta_tables(timesteps, mode="cosine", beta_start=0.0001, beta_end=0.02):
    """models/diffcsp/scheduler.py:54-88 — index 0 is the t=0 pad (beta=0)."""
    if mode == "cosine":
        betas = cosine_beta_schedule(timesteps)
    elif mode == "linear":
        betas = torch.linspace(beta_start, beta_end, timesteps)
    elif mode == "quadratic":
        betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, timesteps) ** 2
    elif mode == "sigmoid":
        betas = torch.sigmoid(torch.linspace(-6, 6, timesteps)) * (beta_end - beta_start) + beta_start
    else:
        raise ValueError(mode)
    betas = torch.cat([torch.zeros([1]), betas], dim=0)
    alphas = 1. - betas
    alphas_cumprod = torch.cumprod(alphas, 0)
    sigmas = torch.zeros_like(betas)
    sigmas[1:] = betas[1:] * (1. - alphas_cumprod[:-1]) / (1. - alphas_cumprod[1:])
    sigmas = torch.sqrt(sigmas)
    return dict(betas=betas, alphas=alphas, alphas_cumprod=alphas_cumprod, sigmas=sigmas)


def p_wrapped_normal(x, sigma, N=10, T=1.0):
    """models/diffcsp/scheduler.py:32-36."""
    p_ = 0
    for i in range(-N, N + 1):
        p_ += torch.exp(-(x + T * i) ** 2 / 2 / sigma ** 2)
    return p_


def d_log_p_wrapped_normal(x, sigma, N=10, T=1.0):
    """models/diffcsp/scheduler.py:39-43 (this is MINUS the score of the wrapped normal)."""
    p_ = 0
    for i in range(-N, N + 1):
        p_ += (x + T * i) / sigma ** 2 * torch.exp(-(x + T * i) ** 2 / 2 / sigma ** 2)
    return p_ / p_wrapped_normal(x, sigma, N, T)


def sigma_norm_mc(sigma, T=1.0, sn=10000, randn_like=torch.randn_like):
    """models/diffcsp/scheduler.py:46-51 — Monte-Carlo, RNG dependent."""
    sigmas = sigma[None, :].repeat(sn, 1)
    x_sample = sigma * randn_like(sigmas)
    x_sample = x_sample % T
    normal_ = d_log_p_wrapped_normal(x_sample, sigmas, T=T)
    return (normal_ ** 2).mean(dim=0)


def sigma_tables(timesteps, sigma_begin=0.01, sigma_end=1.0, sigmas_norm=None):
    """models/diffcsp/scheduler.py:95-112.  `sigmas_norm` (length T+1) may be supplied to copy a
    checkpoint's / another instance's Monte-Carlo buffer instead of redrawing it."""
    sig = torch.FloatTensor(np.exp(np.linspace(np.log(sigma_begin), np.log(sigma_end), timesteps)))
    if sigmas_norm is None:
        sn = torch.cat([torch.ones([1]), sigma_norm_mc(sig)], dim=0)
    else:
        sn = torch.as_tensor(sigmas_norm, dtype=torch.float32).clone()
    return dict(sigmas=torch.cat([torch.zeros([1]), sig], dim=0), sigmas_norm=sn,
                sigma_begin=sigma_begin, sigma_end=sigma_end)


# ----------------------------------------------------------------------------- small pieces
def time_embedding(time, dim):
    """models/diffcsp/diffusion.py:59-66."""
    half = dim // 2
    e = math.log(10000) / (half - 1)
    e = torch.exp(torch.arange(half) * -e)
    e = time[:, None] * e[None, :]
    return torch.cat((e.sin(), e.cos()), dim=-1)


def sinusoids_embedding(x, n_frequencies):
    """models/diffcsp/cspnet.py:12-24 — x [E,3] -> [E, 6F] = [sin(2 pi k x_c)]_{c,k} ‖ [cos]."""
    freq = 2 * math.pi * torch.arange(n_frequencies)
    emb = x.unsqueeze(-1) * freq[None, None, :].to(x.dtype)
    emb = emb.reshape(-1, n_frequencies * 3)
    return torch.cat((emb.sin(), emb.cos()), dim=-1)


def lattice_params_to_matrix(lengths, angles):
    """models/diffcsp/utils.py:68-96."""
    ar = torch.deg2rad(angles)
    c, s = torch.cos(ar), torch.sin(ar)
    val = (c[:, 0] * c[:, 1] - c[:, 2]) / (s[:, 0] * s[:, 1])
    val = torch.clamp(val, -1., 1.)
    gs = torch.arccos(val)
    z = torch.zeros(lengths.size(0), dtype=lengths.dtype)
    va = torch.stack([lengths[:, 0] * s[:, 1], z, lengths[:, 0] * c[:, 1]], dim=1)
    vb = torch.stack([-lengths[:, 1] * s[:, 0] * torch.cos(gs), lengths[:, 1] * s[:, 0] * torch.sin(gs),
                      lengths[:, 1] * c[:, 0]], dim=1)
    vc = torch.stack([z, z, lengths[:, 2]], dim=1)
    return torch.stack([va, vb, vc], dim=1)


def lattices_to_params_shape(lattices):
    """models/diffcsp/sample.py:103-114."""
    lengths = torch.sqrt(torch.sum(lattices ** 2, dim=-1))
    angles = torch.zeros_like(lengths)
    for i in range(3):
        j, k = (i + 1) % 3, (i + 2) % 3
        angles[..., i] = torch.clamp(torch.sum(lattices[..., j, :] * lattices[..., k, :], dim=-1)
                                     / (lengths[..., j] * lengths[..., k]), -1., 1.)
    angles = torch.arccos(angles) * 180.0 / np.pi
    return lengths, angles


def scatter_mean(src, index, dim_size):
    """torch_scatter.scatter(reduce='mean') leaf semantics: sum / clamp(count, 1)."""
    res = torch.zeros((dim_size,) + tuple(src.shape[1:]), dtype=src.dtype).index_add_(0, index, src)
    cnt = torch.zeros(dim_size, dtype=src.dtype).index_add_(0, index, torch.ones_like(index, dtype=src.dtype))
    return res / cnt.clamp_(min=1).view((-1,) + (1,) * (src.dim() - 1))


# ----------------------------------------------------------------------------- edges
def fc_edges(num_atoms):
    """models/diffcsp/cspnet.py:238-242 — all ordered pairs (i,j) incl. i==j inside each crystal,
    row-major (sorted by i then j); block_diag + nonzero restated arithmetically."""
    src, dst = [], []
    off = 0
    for n in [int(v) for v in num_atoms]:
        idx = torch.arange(n)
        src.append((idx[:, None].expand(n, n) + off).reshape(-1))
        dst.append((idx[None, :].expand(n, n) + off).reshape(-1))
        off += n
    if not src:
        return torch.zeros(2, 0, dtype=torch.long)
    return torch.stack([torch.cat(src), torch.cat(dst)])


def repeat_blocks(sizes, repeats, continuous_indexing=True, start_idx=0, block_inc=0, repeat_inc=0):
    """models/diffcsp/utils.py:190-332, restated as the explicit double loop its docstring describes
    (scalar `repeats`/`block_inc`/`repeat_inc` or per-block sequences)."""
    sizes = [int(s) for s in sizes]
    nb = len(sizes)

    def per_block(v, n):
        if isinstance(v, (int, float)):
            return [int(v)] * n
        return [int(t) for t in v]

    reps = per_block(repeats, nb)
    rinc = per_block(repeat_inc, nb)
    binc = per_block(block_inc, max(nb - 1, 0))
    out = []
    base = start_idx
    for b in range(nb):
        for r in range(reps[b]):
            out.extend(base + r * rinc[b] + k for k in range(sizes[b]))
        if continuous_indexing:
            base += sizes[b]
        if b < nb - 1:
            base += binc[b]
    return torch.tensor(out, dtype=torch.long)


def radius_graph_pbc(pos, lattices, natoms, max_neighbors):
    """models/diffcsp/utils.py:335-514 with get_max_neighbors_mask (:517-601), restated per crystal.

    Candidates: all intra-crystal ordered pairs (i1, i2) x 27 image offsets in (-1,0,1)^3 (first
    axis slowest), enumeration order (crystal, i1, i2, cell).  The `radius` argument of the reference
    is ignored (:463): radius = min inter-plane distance + 0.01; keep d2 <= radius^2 and d2 > 1e-4.
    Neighbour cap (:556-579): per centre atom i1, threshold = (K+1)-th smallest d2 (0-based column K
    of the sorted row, +inf when it has <= K candidates) + 0.01; keep d2 < threshold.
    Returns edge_index = (index2, index1) [2,E], cell offsets [E,3] (float), edges per crystal [B].
    """
    natoms = [int(v) for v in natoms]
    cells = torch.tensor([[a, b, c] for a in (-1, 0, 1) for b in (-1, 0, 1) for c in (-1, 0, 1)],
                         dtype=pos.dtype)  # meshgrid order, :417-424
    i1_all, i2_all, uc_all, nb_img = [], [], [], []
    off = 0
    for b, n in enumerate(natoms):
        cell = lattices[b]
        c23 = torch.cross(cell[1], cell[2], dim=-1)
        vol = torch.sum(cell[0] * c23, dim=-1, keepdim=True)
        d1 = 1 / torch.norm(c23 / vol, p=2, dim=-1)
        c31 = torch.cross(cell[2], cell[0], dim=-1)
        d2 = 1 / torch.norm(c31 / vol, p=2, dim=-1)
        c12 = torch.cross(cell[0], cell[1], dim=-1)
        d3 = 1 / torch.norm(c12 / vol, p=2, dim=-1)
        radius = torch.stack([d1, d2, d3]).min() + 0.01
        offs = (cell.t() @ cells.t())  # [3,27]  (:427-431)
        p = pos[off:off + n]
        i1 = torch.arange(n).repeat_interleave(n)
        i2 = torch.arange(n).repeat(n)
        pos1 = p[i1].view(-1, 3, 1).expand(-1, -1, 27)
        pos2 = p[i2].view(-1, 3, 1).expand(-1, -1, 27) + offs.view(1, 3, 27)
        dsq = torch.sum((pos1 - pos2) ** 2, dim=1).view(-1)
        i1e = i1.view(-1, 1).repeat(1, 27).view(-1)
        i2e = i2.view(-1, 1).repeat(1, 27).view(-1)
        uce = cells.view(1, 27, 3).repeat(n * n, 1, 1).view(-1, 3)
        m = torch.logical_and(dsq <= radius * radius, dsq > 0.0001)
        i1e, i2e, uce, dsq = i1e[m], i2e[m], uce[m], dsq[m]
        keep = torch.ones_like(i1e, dtype=torch.bool)
        for a in range(n):
            sel = (i1e == a).nonzero().view(-1)
            if sel.numel() > max_neighbors:
                srt = torch.sort(dsq[sel])[0]
                thr = srt[max_neighbors] + 0.01
                keep[sel] = dsq[sel] < thr
        i1_all.append(i1e[keep] + off)
        i2_all.append(i2e[keep] + off)
        uc_all.append(uce[keep])
        nb_img.append(int(keep.sum()))
        off += n
    index1, index2 = torch.cat(i1_all), torch.cat(i2_all)
    return torch.stack((index2, index1)), torch.cat(uc_all), torch.tensor(nb_img, dtype=torch.long)


def reorder_symmetric_edges(edge_index, cell_offsets, neighbors, edge_vector):
    """models/diffcsp/cspnet.py:159-234: keep (a<b) edges and same-atom edges whose cell is
    lexicographically negative, append their reversals (offsets / vectors sign-flipped), and order
    per crystal as [kept..., reversed...]."""
    co = cell_offsets
    earlier = (co[:, 0] < 0) | ((co[:, 0] == 0) & (co[:, 1] < 0)) | \
              ((co[:, 0] == 0) & (co[:, 1] == 0) & (co[:, 2] < 0))
    mask = (edge_index[0] < edge_index[1]) | ((edge_index[0] == edge_index[1]) & earlier)
    crystal = torch.repeat_interleave(torch.arange(neighbors.numel()), neighbors)
    ei, off, vec, nn = [], [], [], []
    for b in range(neighbors.numel()):
        sel = (mask & (crystal == b)).nonzero().view(-1)
        e = edge_index[:, sel]
        ei += [e, torch.stack([e[1], e[0]])]
        off += [co[sel], -co[sel]]
        vec += [edge_vector[sel], -edge_vector[sel]]
        nn.append(2 * sel.numel())
    return torch.cat(ei, dim=1), torch.cat(off), torch.tensor(nn, dtype=torch.long), torch.cat(vec)


def gen_edges(hp, num_atoms, frac_coords, lattices, node2graph):
    """models/diffcsp/cspnet.py:236-257."""
    if hp["edge_style"] == "fc":
        e = fc_edges(num_atoms)
        return e, (frac_coords[e[1]] - frac_coords[e[0]]) % 1.
    cart = torch.einsum('bi,bij->bj', frac_coords, lattices[node2graph])
    edge_index, to_j, nb = radius_graph_pbc(cart, lattices, num_atoms, hp["max_neighbors"])
    j_index, i_index = edge_index
    dv = frac_coords[j_index] - frac_coords[i_index]
    dv = dv + to_j.to(dv.dtype)
    e_new, _, _, v_new = reorder_symmetric_edges(edge_index, to_j, nb, dv)
    return e_new, -v_new


# ----------------------------------------------------------------------------- score network
def init_params(hp, seed=0, head_scale=0.05, dtype=torch.float32):
    """Seeded default nn.Linear / nn.LayerNorm init in the reference's module-construction order
    (models/diffcsp/cspnet.py:96-147), output heads scaled by `head_scale` (keeps
    random-init trajectories finite).  Keys = the reference decoder's state_dict names."""
    import torch.nn as nn
    g = torch.random.get_rng_state()
    torch.manual_seed(seed)
    H, A = hp["hidden_dim"], hp["max_atoms"]
    dis_dim = hp["num_freqs"] * 6
    mods = [("node_embedding", nn.Linear(A, H)),
            ("atom_latent_emb", nn.Linear(H + hp["latent_dim"] + hp["time_dim"], H))]
    for i in range(hp["num_layers"]):
        p = "csp_layer_%d." % i
        mods += [(p + "edge_mlp.0", nn.Linear(2 * H + 9 + dis_dim, H)), (p + "edge_mlp.2", nn.Linear(H, H)),
                 (p + "node_mlp.0", nn.Linear(2 * H, H)), (p + "node_mlp.2", nn.Linear(H, H))]
        if hp["ln"]:
            mods.append((p + "layer_norm", nn.LayerNorm(H)))
    mods += [("coord_out", nn.Linear(H, 3, bias=False)), ("lattice_out", nn.Linear(H, 9, bias=False))]
    if hp["ln"]:
        mods.append(("final_layer_norm", nn.LayerNorm(H)))
    mods.append(("type_out", nn.Linear(H, MAX_ATOMIC_NUM)))
    sd = {}
    for name, m in mods:
        for k, v in m.state_dict().items():
            sd[name + "." + k] = v.detach().clone()
    for k in ("coord_out.weight", "lattice_out.weight", "type_out.weight", "type_out.bias"):
        sd[k] *= head_scale
    torch.random.set_rng_state(g)
    return {k: v.to(dtype) for k, v in sd.items()}


def _linear(x, sd, name):
    return torch.nn.functional.linear(x, sd[name + ".weight"], sd.get(name + ".bias"))


def _silu(x):
    return torch.nn.functional.silu(x)


def csp_layer(sd, p, hp, h, lattices, edges, edge2graph, frac_diff):
    """models/diffcsp/cspnet.py:59-91."""
    H = hp["hidden_dim"]
    h_in = h
    if hp["ln"]:
        h = torch.nn.functional.layer_norm(h, (H,), sd[p + "layer_norm.weight"], sd[p + "layer_norm.bias"])
    hi, hj = h[edges[0]], h[edges[1]]
    fd = sinusoids_embedding(frac_diff, hp["num_freqs"])
    ips = (lattices @ lattices.transpose(-1, -2)).view(-1, 9)[edge2graph]   # CSPLayer.ip is hard True (:42)
    e_in = torch.cat([hi, hj, ips, fd], dim=1)
    ef = _silu(_linear(_silu(_linear(e_in, sd, p + "edge_mlp.0")), sd, p + "edge_mlp.2"))
    agg = scatter_mean(ef, edges[0], h.shape[0])
    out = _silu(_linear(_silu(_linear(torch.cat([h, agg], dim=1), sd, p + "node_mlp.0")), sd, p + "node_mlp.2"))
    return h_in + out


def cspnet_forward(sd, hp, t, atom_types, frac_coords, lattices, num_atoms, node2graph):
    """models/diffcsp/cspnet.py:260-294 (smooth=True, pred_type=True).  Returns (lattice_out [B,3,3],
    coord_out [N,3], type_out [N,100])."""
    edges, frac_diff = gen_edges(hp, num_atoms, frac_coords, lattices, node2graph)
    edge2graph = node2graph[edges[0]]
    h = _linear(atom_types, sd, "node_embedding")
    h = torch.cat([h, t.repeat_interleave(num_atoms, dim=0)], dim=1)
    h = _linear(h, sd, "atom_latent_emb")
    for i in range(hp["num_layers"]):
        h = csp_layer(sd, "csp_layer_%d." % i, hp, h, lattices, edges, edge2graph, frac_diff)
    if hp["ln"]:
        h = torch.nn.functional.layer_norm(h, (hp["hidden_dim"],), sd["final_layer_norm.weight"],
                                           sd["final_layer_norm.bias"])
    coord_out = _linear(h, sd, "coord_out")
    g = scatter_mean(h, node2graph, int(num_atoms.numel()))
    lat = _linear(g, sd, "lattice_out").view(-1, 3, 3)
    if hp["ip"]:
        lat = torch.einsum('bij,bjk->bik', lat, lattices)
    return lat, coord_out, _linear(h, sd, "type_out")


# ----------------------------------------------------------------------------- diffusion module
class Noise:
    """Draw-order-faithful noise source.  `rand`, `randn` take a shape; default = torch global RNG."""

    def __init__(self, generator=None, dtype=torch.float32, draw_dtype=None):
        self.g, self.dtype, self.draw_dtype = generator, dtype, draw_dtype or dtype

    def rand(self, shape):
        return torch.rand(tuple(shape), generator=self.g, dtype=self.draw_dtype).to(self.dtype)

    def randn(self, shape):
        return torch.randn(tuple(shape), generator=self.g, dtype=self.draw_dtype).to(self.dtype)


class Schedules:
    def __init__(self, hp, sigmas_norm=None):
        self.T = hp["timesteps"]
        self.beta = beta_tables(self.T, hp["beta_mode"])
        self.sigma = sigma_tables(self.T, hp["sigma_begin"], hp["sigma_end"], sigmas_norm)


def add_noise(hp, sch, batch, t_idx, noise):
    """models/diffcsp/diffusion.py:81-119 with an integer time index (times = T - t_idx, :86-87).
    `batch` has lengths, angles [B,3], frac_coords [N,3], atom_types [N] (1-based), num_atoms, batch.
    Draw order: rand_l, rand_x (:102), rand_t (:111)."""
    B = int(batch.num_atoms.numel())
    time_arr = np.arange(sch.T, 0, -1)
    times = torch.full((B,), int(time_arr[t_idx]))
    dt = batch.frac_coords.dtype
    temb = time_embedding(times, hp["time_dim"]).to(dt)
    ac = sch.beta["alphas_cumprod"][times].to(dt)
    c0, c1 = torch.sqrt(ac), torch.sqrt(1. - ac)
    sig = sch.sigma["sigmas"][times].to(dt)
    sn = sch.sigma["sigmas_norm"][times].to(dt)
    lattices = lattice_params_to_matrix(batch.lengths, batch.angles)
    x0 = batch.frac_coords
    rand_l, rand_x = noise.randn(lattices.shape), noise.randn(x0.shape)
    l_t = c0[:, None, None] * lattices + c1[:, None, None] * rand_l
    sig_a = sig.repeat_interleave(batch.num_atoms)[:, None]
    sn_a = sn.repeat_interleave(batch.num_atoms)[:, None]
    x_t = (x0 + sig_a * rand_x) % 1.
    onehot = torch.nn.functional.one_hot(batch.atom_types - 1, num_classes=MAX_ATOMIC_NUM).to(dt)
    rand_t = noise.randn(onehot.shape)
    a_t = c0.repeat_interleave(batch.num_atoms)[:, None] * onehot + \
        c1.repeat_interleave(batch.num_atoms)[:, None] * rand_t
    tar_x = d_log_p_wrapped_normal(sig_a * rand_x, sig_a) / torch.sqrt(sn_a)
    return (temb, a_t, x_t, l_t, batch.num_atoms, batch.batch), (rand_l, tar_x, rand_t), batch.batch


def calc_sample_loss(sd, hp, input_all):
    """models/diffcsp/diffusion.py:121-138 — per-crystal loss [B] and the predictions."""
    noised, (rand_l, tar_x, rand_t), bidx = input_all
    pred_l, pred_x, pred_t = cspnet_forward(sd, hp, *noised)
    B = pred_l.shape[0]
    loss_l = torch.pow(pred_l - rand_l, 2).mean(dim=(1, 2))
    loss_x = scatter_mean(torch.pow(pred_x - tar_x, 2).mean(dim=1), bidx, B)
    loss_t = scatter_mean(torch.pow(pred_t - rand_t, 2).mean(dim=1), bidx, B)
    loss = hp["cost_lattice"] * loss_l + hp["cost_coord"] * loss_x + hp["cost_type"] * loss_t
    return loss, (pred_l, pred_x, pred_t)


def calc_kl_reg(agent_pred, prior_pred, bidx):
    """models/diffcsp/diffusion.py:140-149 — unweighted sum of the three per-crystal MSEs."""
    (pl, px, pt), (ql, qx, qt) = agent_pred, prior_pred
    B = pl.shape[0]
    k0 = torch.pow(pl - ql.detach(), 2).mean(dim=(1, 2))
    k1 = scatter_mean(torch.pow(px - qx.detach(), 2).mean(dim=1), bidx, B)
    k2 = scatter_mean(torch.pow(pt - qt.detach(), 2).mean(dim=1), bidx, B)
    return k0 + k1 + k2


def reverse_step_coeffs(sch, t, step_lr):
    """Scalars of one reverse step, models/diffcsp/diffusion.py:300-307,324-325,341-343 (fp32 tensor
    arithmetic, same op order)."""
    al = sch.beta["alphas"][t]
    ac = sch.beta["alphas_cumprod"][t]
    c0 = 1.0 / torch.sqrt(al)
    c1 = (1 - al) / torch.sqrt(1 - ac)
    sigmas = sch.beta["sigmas"][t]
    sx = sch.sigma["sigmas"][t]
    sn = sch.sigma["sigmas_norm"][t]
    step_c = step_lr * (sx / sch.sigma["sigma_begin"]) ** 2
    std_c = torch.sqrt(2 * step_c)
    adj = sch.sigma["sigmas"][t - 1]
    step_p = (sx ** 2 - adj ** 2)
    std_p = torch.sqrt((adj ** 2 * (sx ** 2 - adj ** 2)) / (sx ** 2))
    return dict(c0=c0, c1=c1, sigmas=sigmas, sqrt_sn=torch.sqrt(sn), step_c=step_c, std_c=std_c,
                step_p=step_p, std_p=std_p)


def sample(sd, hp, sch, num_atoms, noise, step_lr=1e-5, timesteps=None, return_traj=False):
    """models/diffcsp/diffusion.py:273-399, keep_coords = keep_lattice = False.  The four unused
    log-probabilities (:353-382) are not computed; the RNG draw ORDER is kept: x_T, l_T, t_T (:277-279)
    then per step rand_l, rand_t, rand_x twice (:320-322, 337-339), zeros at t == 1.
    `timesteps` < T runs only the LAST `timesteps` steps' worth of loop from t = timesteps (used for
    bounded CPU samples; schedule tables still indexed by t)."""
    num_atoms = torch.as_tensor(num_atoms, dtype=torch.long)
    B, N = int(num_atoms.numel()), int(num_atoms.sum())
    n2g = torch.repeat_interleave(torch.arange(B), num_atoms)
    dt = sd["coord_out.weight"].dtype
    x_T = noise.rand([N, 3])
    l_T = noise.randn([B, 3, 3])
    t_T = noise.randn([N, MAX_ATOMIC_NUM])
    T0 = sch.T if timesteps is None else timesteps
    x_t, l_t, a_t = x_T % 1., l_T, t_T
    traj = []
    for t in range(T0, 0, -1):
        temb = time_embedding(torch.full((B,), t), hp["time_dim"]).to(dt)
        c = {k: v.to(dt) for k, v in reverse_step_coeffs(sch, t, step_lr).items()}

        def draw():
            if t > 1:
                return noise.randn(l_T.shape), noise.randn(t_T.shape), noise.randn(x_T.shape)
            return torch.zeros_like(l_T), torch.zeros_like(t_T), torch.zeros_like(x_T)
        # corrector (:320-334)
        _, _, rand_x = draw()
        pred_l, pred_x, pred_t = cspnet_forward(sd, hp, temb, a_t, x_t, l_t, num_atoms, n2g)
        pred_x = pred_x * c["sqrt_sn"]
        x_half = x_t - c["step_c"] * pred_x + c["std_c"] * rand_x
        # predictor (:337-351)
        rand_l, rand_t, rand_x = draw()
        pred_l, pred_x, pred_t = cspnet_forward(sd, hp, temb, a_t, x_half, l_t, num_atoms, n2g)
        pred_x = pred_x * c["sqrt_sn"]
        x_next = x_half - c["step_p"] * pred_x + c["std_p"] * rand_x
        l_t = c["c0"] * (l_t - c["c1"] * pred_l) + c["sigmas"] * rand_l
        a_t = c["c0"] * (a_t - c["c1"] * pred_t) + c["sigmas"] * rand_t
        x_t = (x_next % 1.) % 1.      # :351 then :386
        if return_traj:
            traj.append(dict(frac_coords=x_t.clone(), lattices=l_t.clone(), atom_types=a_t.clone()))
    out = dict(frac_coords=x_t, lattices=l_t, atom_types=a_t, num_atoms=num_atoms, batch_idx=n2g)
    return (out, traj) if return_traj else out


def generate_postprocess(out):
    """models/diffcsp/sample.py:174-199: argmax atom types (+1), lattice -> lengths/angles, split."""
    lengths, angles = lattices_to_params_shape(out["lattices"])
    types = torch.argmax(out["atom_types"], dim=-1) + 1
    off = [0] + torch.cumsum(out["num_atoms"], 0).tolist()
    crystals = []
    for i in range(len(off) - 1):
        crystals.append(dict(frac_coords=out["frac_coords"][off[i]:off[i + 1]],
                             atom_types=types[off[i]:off[i + 1]], lengths=lengths[i].view(1, -1),
                             angles=angles[i].view(1, -1), num_atoms=int(out["num_atoms"][i])))
    return crystals


# ----------------------------------------------------------------------------- fine-tune step
def ft_timestep_loss(agent_sd, prior_sd, hp, sch, batch, reward, t_idx, noise, sigma, accum_steps):
    """One inner iteration of MatInvent.ft_step (pipeline/mat_invent.py:152-163): returns the scalar
    that is back-propagated plus the per-crystal pieces."""
    noised = add_noise(hp, sch, batch, t_idx, noise)
    sample_loss, agent_pred = calc_sample_loss(agent_sd, hp, noised)
    with torch.no_grad():
        _, prior_pred = calc_sample_loss(prior_sd, hp, noised)
    loss_diff = reward * sample_loss
    kl = calc_kl_reg(agent_pred, prior_pred, batch.batch)
    loss_kl = kl * (1.1 - reward)
    loss = (loss_diff + loss_kl * sigma).mean() / accum_steps
    return loss, dict(sample_loss=sample_loss, kl=kl, loss_diff=loss_diff, loss_kl=loss_kl,
                      agent_pred=agent_pred, prior_pred=prior_pred, noised=noised)


def ft_step(agent_sd, prior_sd, hp, sch, batch, reward, noise, lr=1e-4, accum_steps=50, epochs=3,
            sigma=0.025, timesteps=None):
    """pipeline/mat_invent.py:125-189 on ONE batch: fresh Adam (torch defaults, :136), for each epoch
    for t in range(timesteps): backward of the reward-weighted loss, Adam step every `accum_steps`.
    Mutates and returns `agent_sd`; also returns the per-epoch (loss, loss_diff, loss_kl) logs."""
    timesteps = sch.T if timesteps is None else timesteps
    params = [agent_sd[k].requires_grad_(True) for k in agent_sd]
    opt = torch.optim.Adam(params, lr=lr)
    logs = []
    B = int(batch.num_atoms.numel())
    for _ in range(epochs):
        opt.zero_grad()
        loss_s = diff_s = kl_s = 0.
        for t in range(timesteps):
            loss, parts = ft_timestep_loss(agent_sd, prior_sd, hp, sch, batch, reward, t, noise, sigma,
                                           accum_steps)
            loss.backward()
            if (t + 1) % accum_steps == 0:
                opt.step()
                opt.zero_grad()
            loss_s += loss.item() * accum_steps
            diff_s += parts["loss_diff"].sum().item()
            kl_s += parts["loss_kl"].sum().item()
        if timesteps % accum_steps != 0:
            opt.step()
        logs.append((loss_s / timesteps * B / B, diff_s / timesteps / B, kl_s / timesteps / B))
    return agent_sd, logs


# ----------------------------------------------------------------------------- replay buffer
def reduced_composition_key(atom_types):
    """Equivalence class used by ReplayBuffer dedupe: pymatgen `composition.reduced_formula`
    (memory/replay_buffer.py:38) == the gcd-reduced element-count vector."""
    z = [int(v) for v in atom_types]
    cnt = {}
    for v in z:
        cnt[v] = cnt.get(v, 0) + 1
    g = 0
    for v in cnt.values():
        g = math.gcd(g, v)
    return tuple(sorted((k, v // g) for k, v in cnt.items()))


class ReplayBufferOracle:
    """memory/replay_buffer.py:11-104 restated on plain lists (no pandas).  Ties in reward keep the
    earlier row (a stable sort; pandas' quicksort leaves tie order unspecified).  PINNED: row-for-row equal to the
    unmodified reference class after every extend / sample / memory_purge of a random call sequence
    (tests/test_oracle_vs_reference.py::test_replay_buffer_oracle_matches_live_reference); `sample(np.random)`
    consumes numpy's global RNG exactly like `DataFrame.sample` does."""

    def __init__(self, buffer_size=100, sample_size=8, reward_cutoff=0.0):
        self.buffer_size, self.sample_size, self.reward_cutoff = buffer_size, sample_size, reward_cutoff
        self.rows = []   # (data, key, reward)

    def extend(self, data, keys, rewards):
        rows = self.rows + [(d, k, float(r)) for d, k, r in zip(data, keys, rewards)]
        rows.sort(key=lambda r: -r[2])
        seen, uniq = set(), []
        for r in rows:
            if r[1] not in seen:
                seen.add(r[1])
                uniq.append(r)
        self.rows = [r for r in uniq[:self.buffer_size] if r[2] > self.reward_cutoff]

    def sample(self, rng):
        k = min(len(self.rows), self.sample_size)
        if k == 0:
            return [], []
        idx = rng.choice(len(self.rows), k, replace=False)
        return [self.rows[i][0] for i in idx], np.array([self.rows[i][2] for i in idx])

    def memory_purge(self, keys):
        ks = set(keys)
        self.rows = [r for r in self.rows if r[1] not in ks]

    def __len__(self):
        return len(self.rows)
