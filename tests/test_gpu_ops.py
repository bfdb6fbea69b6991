"""GPU: every C-ABI operator against a float64 torch statement of the same op (tolerances written per test)."""
import math

import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from matinvent_b200 import ops as o
    return o


def _rand(*s, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*s, generator=g).cuda()


@pytest.mark.parametrize("tA,tB", [(False, True), (False, False), (True, False), (True, True)])
@pytest.mark.parametrize("M,N,K", [(1, 1, 1), (5, 3, 9), (130, 100, 100), (257, 512, 768), (64, 9, 512), (300, 1024, 512)])
def test_sgemm_layouts(ops, tA, tB, M, N, K):
    A = _rand(K, M, seed=1) if tA else _rand(M, K, seed=1)
    B = _rand(N, K, seed=2) if tB else _rand(K, N, seed=2)
    C = torch.full((M, N), float("nan"), device="cuda")
    ops.sgemm(A, B, C, transA=tA, transB=tB)
    ref = (A.double().t() if tA else A.double()) @ (B.double().t() if tB else B.double())
    assert rel_err(C, ref) < 2e-6      # fp32 accumulate over K <= 768


def test_sgemm_strided_views_and_unaligned(ops):
    big = _rand(200, 1024, seed=3)
    A = big[:, :512]                    # lda = 1024
    W = _rand(384, 1801, seed=4)[:, 1033:1033 + 512]   # unaligned rows (ld 1801, offset 1033): scalar path
    C = torch.zeros(200, 768, device="cuda")[:, 384:]  # ldc = 768
    ops.sgemm(A, W, C)
    assert rel_err(C, A.double() @ W.double().t()) < 2e-6


def test_sgemm_fused_epilogue(ops):
    M, N, K = 333, 512, 96
    A, W, bias = _rand(M, K, seed=5), _rand(N, K, seed=6), _rand(N, seed=7)
    P, Q, Cb = _rand(40, 2 * N, seed=8), _rand(40, N, seed=9), _rand(7, N, seed=10)
    g = torch.Generator().manual_seed(11)
    i1 = torch.randint(0, 40, (M,), generator=g).int().cuda()
    i2 = torch.randint(0, 40, (M,), generator=g).int().cuda()
    i3 = torch.randint(0, 7, (M,), generator=g).int().cuda()
    R = _rand(M, N, seed=12)
    Z = torch.empty(M, N, device="cuda")
    C = torch.empty(M, N, device="cuda")
    ops.sgemm(A, W, C, bias=bias, gathers=[(P[:, :N], i1), (Q, i2), (Cb, i3)], z_out=Z, act=ops.ACT_SILU, resid=R)
    z = A.double() @ W.double().t() + bias.double() + P.double()[i1.long(), :N] + Q.double()[i2.long()] + Cb.double()[i3.long()]
    assert rel_err(Z, z) < 2e-6
    assert rel_err(C, torch.nn.functional.silu(z) + R.double()) < 2e-6
    # DSILU epilogue: C = (A W^T) * silu'(Zin)
    C2 = torch.empty(M, N, device="cuda")
    ops.sgemm(A, W, C2, act=ops.ACT_DSILU, z_in=Z)
    zz = Z.double().requires_grad_(True)
    torch.nn.functional.silu(zz).sum().backward()
    assert rel_err(C2, (A.double() @ W.double().t()) * zz.grad) < 2e-6
    # beta accumulate
    C3 = R.clone()
    ops.sgemm(A, W, C3, beta=1.0)
    assert rel_err(C3, A.double() @ W.double().t() + R.double()) < 2e-6


@pytest.mark.parametrize("M,N,K,splitk", [(512, 768, 5000, 8), (9, 512, 37, 4), (100, 512, 4, 2)])
def test_sgemm_splitk_accumulates(ops, M, N, K, splitk):
    dY, X = _rand(K, M, seed=13), _rand(K, N, seed=14)
    G = _rand(M, N, seed=15)
    G0 = G.clone()
    ops.sgemm(dY, X, G, transA=True, transB=False, beta=1.0, splitk=splitk)
    assert rel_err(G, G0.double() + dY.double().t() @ X.double()) < 3e-6


def test_fc_edges_match_reference_order(ops):
    from oracle import diffcsp_oracle as O
    from matinvent_b200.models.diffcsp.graph import CrystalGraph
    na = [3, 1, 7, 20, 5, 12]
    g = CrystalGraph(na, "cuda")
    e = O.fc_edges(na)
    assert torch.equal(g.edge_src.cpu().long(), e[0]) and torch.equal(g.edge_dst.cpu().long(), e[1])
    n2g = torch.repeat_interleave(torch.arange(len(na)), torch.tensor(na))
    assert torch.equal(g.node_graph.cpu().long(), n2g)
    assert torch.equal(g.edge_graph.cpu().long(), n2g[e[0]])
    # CSR over src and dst
    sp, dp, perm = g.seg_ptr.cpu().long(), g.dst_ptr.cpu().long(), g.dst_perm.cpu().long()
    for i in range(g.N):
        assert torch.all(e[0][sp[i]:sp[i + 1]] == i)
        assert torch.all(e[1][perm[dp[i]:dp[i + 1]]] == i)
    assert sorted(perm.tolist()) == list(range(g.E))


def test_edge_fourier_matches_reference_arithmetic(ops):
    from oracle import diffcsp_oracle as O
    from matinvent_b200.models.diffcsp.graph import CrystalGraph
    na = [4, 9, 20]
    g = CrystalGraph(na, "cuda")
    x = torch.rand(g.N, 3, generator=torch.Generator().manual_seed(3))
    x[0, 0], x[1, 0] = 0.25, 0.25 + 1e-9           # exercises (-1e-9) % 1 == 1.0
    F = 128
    fd = torch.empty(g.E, 3, device="cuda")
    phi = torch.empty(g.E, 6 * F, device="cuda")
    ops.edge_fourier(x.cuda(), g.edge_src, g.edge_dst, None, g.E, F, fd, phi)
    e = O.fc_edges(na)
    fd_ref = (x[e[1]] - x[e[0]]) % 1.
    assert torch.equal(fd.cpu(), fd_ref)          # bit exact: same fp32 subtract + remainder
    ref = O.sinusoids_embedding(fd_ref, F)
    # same fp32 argument; sinf/cosf of CUDA vs CPU libm differ by <= 2 ulp of the result
    assert float((phi.cpu() - ref).abs().max()) < 5e-7


def test_segment_reduce_and_gather_backward(ops):
    from matinvent_b200.models.diffcsp.graph import CrystalGraph
    g = CrystalGraph([3, 1, 7, 20, 5], "cuda")
    H = 512
    X = _rand(g.E, H, seed=20)
    out = torch.empty(g.N, H, device="cuda")
    ops.segment_reduce(X, g.seg_ptr, out, g.N, H, mean=True)
    src = g.edge_src.long()
    ref = torch.zeros(g.N, H, dtype=torch.float64, device="cuda").index_add_(0, src, X.double())
    cnt = torch.bincount(src, minlength=g.N).double()
    assert rel_err(out, ref / cnt[:, None]) < 1e-6
    # sum grouped by destination through the permutation
    ops.segment_reduce(X, g.dst_ptr, out, g.N, H, perm=g.dst_perm, mean=False)
    ref = torch.zeros(g.N, H, dtype=torch.float64, device="cuda").index_add_(0, g.edge_dst.long(), X.double())
    assert rel_err(out, ref) < 1e-6
    am = torch.zeros(g.N, device="cuda")
    ops.segment_reduce(X, g.seg_ptr, out, g.N, H, mean=True, amax_out=am)
    assert torch.equal(am, out.abs().amax(dim=1))
    # strided output + accumulate
    cat = torch.ones(g.N, 2 * H, device="cuda")
    ops.segment_reduce(X, g.seg_ptr, cat[:, H:], g.N, H, mean=False, accumulate=True)
    ref = 1 + torch.zeros(g.N, H, dtype=torch.float64, device="cuda").index_add_(0, src, X.double())
    assert rel_err(cat[:, H:], ref) < 1e-6 and float(cat[:, :H].min()) == 1.0
    # backward of mean + silu
    dOut, Z = _rand(g.N, H, seed=21), _rand(g.E, H, seed=22)
    dX = torch.empty(g.E, H, device="cuda")
    ops.gather_rows_dsilu(dOut, g.edge_src, g.seg_ptr, Z, dX, g.E, H)
    z = Z.double().requires_grad_(True)
    a = torch.nn.functional.silu(z)
    agg = torch.zeros(g.N, H, dtype=torch.float64, device="cuda").index_add_(0, src, a) / cnt[:, None]
    (agg * dOut.double()).sum().backward()
    assert rel_err(dX, z.grad) < 1e-6


def test_colsum(ops):
    X = _rand(1000, 100, seed=23)
    out = torch.ones(100, device="cuda")
    ops.colsum(X, 1000, 100, out, accumulate=True)
    assert rel_err(out, 1 + X.double().sum(0)) < 1e-5
    ops.colsum(X, 1000, 100, out, accumulate=False)
    assert rel_err(out, X.double().sum(0)) < 1e-5


def test_layernorm_fwd_bwd(ops):
    rows, H = 77, 512
    cat = _rand(rows, 2 * H, seed=24)
    x = _rand(rows, H, seed=25) * 3 + 1
    gamma, beta = _rand(H, seed=26), _rand(H, seed=27)
    mean, rstd = torch.empty(rows, device="cuda"), torch.empty(rows, device="cuda")
    ops.layernorm_fwd(x, gamma, beta, cat[:, :H], rows, H, mean, rstd)
    xd = x.double().requires_grad_(True)
    gd, bd = gamma.double().requires_grad_(True), beta.double().requires_grad_(True)
    y = torch.nn.functional.layer_norm(xd, (H,), gd, bd, 1e-5)
    assert rel_err(cat[:, :H], y) < 1e-6
    dy = _rand(rows, H, seed=28)
    (y * dy.double()).sum().backward()
    dx = torch.ones(rows, H, device="cuda")
    dg, db = torch.zeros(H, device="cuda"), torch.zeros(H, device="cuda")
    ops.layernorm_bwd(dy, x, gamma, mean, rstd, dx, dg, db, rows, H, accumulate_dx=True)
    assert rel_err(dx, 1 + xd.grad) < 2e-6
    assert rel_err(dg, gd.grad) < 2e-6 and rel_err(db, bd.grad) < 2e-6


def test_small_per_crystal_ops(ops):
    from oracle import diffcsp_oracle as O
    B = 9
    L = _rand(B, 3, 3, seed=29)
    ips = torch.empty(B, 9, device="cuda")
    ops.lattice_ip(L, ips, B)
    assert rel_err(ips, (L.double() @ L.double().transpose(1, 2)).view(B, 9)) < 1e-6
    Wl, bl = _rand(512, 9, seed=33), _rand(512, seed=34)
    cb = torch.empty(B, 512, device="cuda")
    ops.lattice_linear(L, Wl, bl, cb, B, 512)
    assert rel_err(cb, (L.double() @ L.double().transpose(1, 2)).view(B, 9) @ Wl.double().t() + bl.double()) < 1e-6
    # three equally spaced weight sets (the layers of the flat weight buffer) in one launch, strided output
    flat = _rand(3, 6000, seed=35)
    Ws, bs = flat[0, :4608].view(512, 9), flat[0, 4608:5120]
    cbs = torch.zeros(3, B, 1024, device="cuda")
    ops.lattice_linear(L, Ws, bs, cbs[0, :, :512], B, 512, n_sets=3, w_stride=6000, bias_stride=6000, out_stride=cbs.stride(0))
    for k in range(3):
        ref = (L.double() @ L.double().transpose(1, 2)).view(B, 9) @ flat[k, :4608].view(512, 9).double().t() + flat[k, 4608:5120].double()
        assert rel_err(cbs[k, :, :512], ref) < 1e-6
    assert float(cbs[:, :, 512:].abs().max()) == 0.0
    A = _rand(B, 3, 3, seed=30)
    out = torch.empty(B, 3, 3, device="cuda")
    ops.bmm3(A, L, out, B)
    assert rel_err(out, A.double() @ L.double()) < 1e-6
    ops.bmm3(A, L, out, B, transL=True)
    assert rel_err(out, A.double() @ L.double().transpose(1, 2)) < 1e-6
    g = torch.Generator().manual_seed(31)
    lengths, angles = 3 + 5 * torch.rand(B, 3, generator=g), 70 + 40 * torch.rand(B, 3, generator=g)
    M = torch.empty(B, 3, 3, device="cuda")
    ops.lattice_params_to_matrix(lengths.cuda(), angles.cuda(), M, B)
    assert rel_err(M, O.lattice_params_to_matrix(lengths, angles)) < 2e-6
    l2, a2 = torch.empty(B, 3, device="cuda"), torch.empty(B, 3, device="cuda")
    ops.lattice_matrix_to_params(M, l2, a2, B)
    rl, ra = O.lattices_to_params_shape(M.cpu())
    assert rel_err(l2, rl) < 1e-6 and float((a2.cpu() - ra).abs().max()) < 2e-3   # acos near-cancellation, degrees
    assert float((l2.cpu() - lengths).abs().max()) < 1e-4 and float((a2.cpu() - angles).abs().max()) < 5e-3
    # time embedding
    from matinvent_b200.models.diffcsp.scheduler import time_frequencies
    t = torch.tensor([1, 17, 500, 1000], dtype=torch.int32)
    te = torch.empty(4, 256, device="cuda")
    ops.time_embed(t.cuda(), time_frequencies(256).cuda(), 4, 256, te)
    assert float((te.cpu() - O.time_embedding(t.long(), 256)).abs().max()) < 5e-7
    # argmax (+1), first index on ties like torch
    a = _rand(50, 100, seed=32)
    a[3, 7] = a[3, 70] = 99.0
    idx = torch.empty(50, dtype=torch.int32, device="cuda")
    ops.argmax_rows(a, 50, 100, idx, add=1)
    assert torch.equal(idx.cpu().long(), a.cpu().argmax(-1) + 1) and int(idx[3]) == 8


def test_add_noise_and_losses_match_oracle(ops, gold_small):
    from oracle import diffcsp_oracle as O
    from oracle.ref_import import make_batch
    from conftest import build_module
    from matinvent_b200.models.diffcsp import TapeNoise, CrystalBatch
    gs = gold_small
    hp, ft = gs["hp"], gs["ft"]
    m = build_module(hp, gs["sd"], gs["sigmas_norm"])
    cr = ft["crystals"]
    batch = make_batch(ft["num_atoms"].tolist(), **cr)
    noised, (rl, tx, rt), _ = m.add_noise(batch, ft["t_idx"], noise=TapeNoise("cuda", seed=ft["noise_seed"]))
    temb, a_t, x_t, l_t = noised[:4]
    assert float((temb.cpu() - ft["ref_temb"]).abs().max()) < 1e-6
    assert torch.equal(rl.cpu(), ft["ref_rand_l"]) and torch.equal(rt.cpu(), ft["ref_rand_t"])
    assert rel_err(a_t, ft["ref_a_t"]) < 1e-6 and rel_err(l_t, ft["ref_l_t"]) < 1e-6
    assert float((x_t.cpu() - ft["ref_x_t"]).abs().max()) < 1e-6
    assert rel_err(tx, ft["ref_tar_x"]) < 2e-5          # 21-term exp sums, expf CUDA vs CPU
    # loss kernel against the oracle's per-crystal losses on the reference's own predictions
    pred = [p.cuda().contiguous() for p in ft["ref_agent_pred"]]
    prior = [p.cuda().contiguous() for p in ft["ref_prior_pred"]]
    g = m.decoder.graph_for(ft["num_atoms"])
    B = g.B
    loss, kl = torch.empty(B, device="cuda"), torch.empty(B, device="cuda")
    tgt = (ft["ref_rand_l"].cuda(), ft["ref_tar_x"].cuda(), ft["ref_rand_t"].cuda())
    r = cr["reward"].cuda()
    wk = (ft["sigma"] * (1.1 - r)).contiguous()
    d = [torch.empty_like(p) for p in pred]
    scale = 1.0 / (B * ft["accum"])
    ops.rl_loss(pred, tgt, prior, g.node_off, B, 100, m._costs(), r, wk, scale, loss, kl, d)
    assert rel_err(loss, ft["ref_sample_loss"]) < 1e-5 and rel_err(kl, ft["ref_kl"]) < 1e-5
    # gradient of the scalar objective w.r.t. the predictions, by float64 autograd
    pd = [p.double().cpu().requires_grad_(True) for p in ft["ref_agent_pred"]]
    bidx = batch.batch
    sl = hp["cost_lattice"] * (pd[0] - ft["ref_rand_l"].double()).pow(2).mean(dim=(1, 2)) \
        + hp["cost_coord"] * O.scatter_mean((pd[1] - ft["ref_tar_x"].double()).pow(2).mean(1), bidx, B) \
        + hp["cost_type"] * O.scatter_mean((pd[2] - ft["ref_rand_t"].double()).pow(2).mean(1), bidx, B)
    klv = O.calc_kl_reg(pd, [p.double() for p in ft["ref_prior_pred"]], bidx)
    rr = cr["reward"].double()
    J = (rr * sl + ft["sigma"] * (1.1 - rr) * klv).mean() / ft["accum"]
    assert abs(float(J) - float(ft["ref_loss"])) < 1e-6 * max(1.0, abs(float(J)))
    J.backward()
    for k in range(3):
        assert rel_err(d[k], pd[k].grad) < 1e-5


def test_adam_matches_torch(ops):
    n = 10007
    p0, g = _rand(n, seed=40), _rand(n, seed=41) * 0.01
    p = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([p], lr=1e-4)
    mine, m, v = p0.clone(), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    for step in range(1, 6):
        gi = g * step
        p.grad = gi.clone()
        opt.step()
        gbuf = gi.clone()
        ops.adam_step(mine, gbuf, m, v, 1e-4, step)
        assert float(gbuf.abs().max()) == 0.0
    assert float((mine - p.detach()).abs().max()) < 1e-7


def test_philox_statistics_and_offsets(ops):
    n = 1 << 20
    a = torch.empty(n, device="cuda")
    ops.philox_normal(a, seed=7)
    assert abs(float(a.mean())) < 5e-3 and abs(float(a.std()) - 1) < 5e-3
    assert abs(float((a ** 4).mean()) - 3) < 0.05
    b = torch.empty(n, device="cuda")
    off = torch.zeros(1, dtype=torch.int64, device="cuda")
    ops.philox_normal(b, seed=7, offset_dev=off, advance=True)
    assert torch.equal(a, b) and int(off) == n // 4
    ops.philox_normal(b, seed=7, offset_dev=off, advance=True)
    assert not torch.equal(a, b) and int(off) == n // 2
    c = torch.empty(n, device="cuda")
    ops.philox_normal(c, seed=7, offset=n // 4)
    assert torch.equal(b, c)
    u = torch.empty(n, device="cuda")
    ops.philox_uniform(u, seed=3)
    assert float(u.min()) >= 0 and float(u.max()) < 1 and abs(float(u.mean()) - 0.5) < 2e-3




def test_output_heads_fused_kernel(ops):
    """mi_output_heads (final LayerNorm + coordinate / type / lattice heads, one CTA per crystal) against float64 of
    cspnet.py:276-294: crystals of 1..20 atoms and one longer than a shared-memory pass, with and without LayerNorm,
    with and without the lattice product, partial head sets"""
    g = torch.Generator().manual_seed(12)
    ns = torch.randint(1, 21, (37,), generator=g).tolist() + [70, 1]
    off = torch.tensor([0] + torch.cumsum(torch.tensor(ns), 0).tolist(), dtype=torch.int32).cuda()
    N, B, H, A = sum(ns), len(ns), 512, 100
    h = (_rand(N, H, seed=30) * torch.exp(_rand(N, 1, seed=31))).contiguous()
    gam, bet = 1 + 0.2 * _rand(H, seed=32), 0.1 * _rand(H, seed=33)
    cw, tw, tb, lw = _rand(3, H, seed=34) / 20, _rand(A, H, seed=35) / 20, _rand(A, seed=36), _rand(9, H, seed=37) / 20
    L = _rand(B, 3, 3, seed=38)
    seg = torch.repeat_interleave(torch.arange(B), torch.tensor(ns)).cuda()
    for ln in (True, False):
        for ip in (True, False):
            hf = torch.nn.functional.layer_norm(h.double(), (H,), gam.double(), bet.double(), 1e-5) if ln else h.double()
            rx, ra = hf @ cw.double().t(), hf @ tw.double().t() + tb.double()
            gm = torch.zeros(B, H, dtype=torch.float64, device="cuda").index_add_(0, seg, hf) / torch.tensor(ns, device="cuda").double()[:, None]
            rl = (gm @ lw.double().t()).view(B, 3, 3)
            if ip:
                rl = rl @ L.double()
            px, pa, pl = torch.full((N, 3), 7.0, device="cuda"), torch.full((N, A), 7.0, device="cuda"), torch.full((B, 3, 3), 7.0, device="cuda")
            ops.output_heads(h, off, B, H, gam if ln else None, bet if ln else None, cw, px, tw, tb, pa, lw, L, ip, pl)
            assert rel_err(px, rx) < 2e-6 and rel_err(pa, ra) < 2e-6 and rel_err(pl, rl) < 2e-6, (ln, ip)
    # coordinate head alone (the corrector evaluation): the other outputs are not touched
    px2 = torch.empty(N, 3, device="cuda")
    ops.output_heads(h, off, B, H, gam, bet, cw, px2, tw, tb, None, lw, L, True, None)
    hf = torch.nn.functional.layer_norm(h.double(), (H,), gam.double(), bet.double(), 1e-5)
    assert rel_err(px2, hf @ cw.double().t()) < 2e-6
