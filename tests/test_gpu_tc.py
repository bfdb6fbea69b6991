"""GPU: the tcgen05 3xTF32 GEMM (mi_tc_gemm) against float64 and against the FP32 CUDA-core GEMM.
Tolerance: FP32-grade — max error relative to the result's max-norm below 5e-6 for K <= 1024 with random-sign
operands and 1e-5 with all-positive operands (the FFMA kernel measures ~2e-6; plain TF32 would sit at ~1e-3).
The residual is the tensor core's truncating accumulate (grows with the number of accumulating MMAs, K/8);
measured on B200: K=768 2.4e-6, K=1024 3.5e-6, positive operands 6.0e-6."""
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu


def _rand(*s, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*s, generator=g).cuda()


def _split(ops, W):
    hi, lo = torch.empty_like(W, dtype=torch.float16), torch.empty_like(W, dtype=torch.float16)
    ops.f16_split(W, hi, lo)
    return hi, lo


@pytest.mark.parametrize("M,N,K", [(128, 256, 32), (128, 256, 64), (128, 256, 768), (1, 8, 32), (300, 512, 512), (2643, 1024, 512),
                                   (257, 100, 512), (130, 3, 512), (34445, 512, 768), (256, 512, 256), (77, 512, 104), (1000, 512, 1024), (70000, 256, 64)])
def test_tc_gemm_fp32_grade(M, N, K):
    from matinvent_b200 import ops
    A, W = _rand(M, K, seed=1), _rand(N, K, seed=2)
    hi, lo = _split(ops, W)
    assert float((hi.float() + lo.float() / 2048 - W).abs().max()) <= 2.0 ** -22 * float(W.abs().max())
    C = torch.full((M, N), float("nan"), device="cuda")
    ops.tc_gemm(A, hi, lo, C)
    ref = A.double() @ W.double().t()
    err = rel_err(C, ref)
    print("tc_gemm M=%d N=%d K=%d rel err %.2e" % (M, N, K, err))
    assert err < 5e-6, err


def test_tc_gemm_rejects_unaligned_weights():
    """K = 100 -> 200-byte fp16 rows: not TMA-legal; the host mirror routes such GEMMs to mi_sgemm (ops.tc_ok)."""
    from matinvent_b200 import ops
    from matinvent_b200._lib import MatInventLibError
    A, W = _rand(8, 100, seed=1), _rand(16, 100, seed=2)
    assert not ops.tc_ok(A, W)
    hi, lo = _split(ops, W)
    with pytest.raises(MatInventLibError):
        ops.tc_gemm(A, hi, lo, torch.empty(8, 16, device="cuda"))


def test_tc_gemm_fused_epilogue_and_views():
    from matinvent_b200 import ops
    M, N, K = 333, 512, 96
    big = _rand(M, 2 * K, seed=5)
    A = big[:, K:]                                      # strided A view (lda = 2K)
    W, bias = _rand(N, K, seed=6), _rand(N, seed=7)
    hi, lo = _split(ops, W)
    P, Q, Cb = _rand(40, 2 * N, seed=8), _rand(40, N, seed=9), _rand(7, N, seed=10)
    g = torch.Generator().manual_seed(11)
    i1 = torch.randint(0, 40, (M,), generator=g).int().cuda()
    i2 = torch.randint(0, 40, (M,), generator=g).int().cuda()
    i3 = torch.randint(0, 7, (M,), generator=g).int().cuda()
    R = _rand(M, N, seed=12)
    Z = torch.empty(M, N, device="cuda")
    out = torch.zeros(M, 2 * N, device="cuda")
    C = out[:, N:]                                      # strided C view
    ops.tc_gemm(A, hi, lo, C, bias=bias, gathers=[(P[:, :N], i1), (Q, i2), (Cb, i3)], z_out=Z, act=ops.ACT_SILU, resid=R)
    z = A.double() @ W.double().t() + bias.double() + P.double()[i1.long(), :N] + Q.double()[i2.long()] + Cb.double()[i3.long()]
    assert rel_err(Z, z) < 3e-6
    assert rel_err(C, torch.nn.functional.silu(z) + R.double()) < 5e-6     # + ex2/rcp-approx SiLU (~3e-7)
    assert float(out[:, :N].abs().max()) == 0.0


def test_tc_gemm_matches_ffma_path_statistically():
    """error vs float64 of the two kernels on the edge-GEMM shape: same order of magnitude"""
    from matinvent_b200 import ops
    M, N, K = 4096, 512, 768
    A, W = _rand(M, K, seed=21).abs(), _rand(N, K, seed=22).abs()      # all-positive: worst case for truncation bias
    hi, lo = _split(ops, W)
    C1, C2 = torch.empty(M, N, device="cuda"), torch.empty(M, N, device="cuda")
    ops.tc_gemm(A, hi, lo, C1)
    ops.sgemm(A, W, C2)
    ref = A.double() @ W.double().t()
    e1, e2 = rel_err(C1, ref), rel_err(C2, ref)
    print("positive operands: tc %.2e  ffma %.2e" % (e1, e2))
    assert e1 < 1e-5 and e2 < 3e-6, (e1, e2)


def test_tc_gemm_row_rescaling_keeps_fp32_range():
    """Rows far outside fp16 range (|a| up to 1e9, as in random-init reverse diffusion) with the producer's row
    maxima: same relative accuracy as in-range rows; without the maxima the fp16 split overflows."""
    from matinvent_b200 import ops
    M, N, K = 300, 512, 512
    A, W = _rand(M, K, seed=31), _rand(N, K, seed=32)
    scale = torch.logspace(-6, 9, M).cuda()[:, None]
    A = A * scale
    hi, lo = _split(ops, W)
    amax = A.abs().amax(dim=1).contiguous()
    C = torch.empty(M, N, device="cuda")
    out_amax = torch.zeros(M, device="cuda")
    ops.tc_gemm(A, hi, lo, C, a_amax=amax, amax_out=out_amax)
    ref = A.double() @ W.double().t()
    row_err = ((C.double() - ref).abs().amax(dim=1) / ref.abs().amax(dim=1))
    assert float(row_err.max()) < 5e-6, float(row_err.max())
    assert torch.equal(out_amax, C.abs().amax(dim=1))
    C2 = torch.empty(M, N, device="cuda")
    ops.tc_gemm(A, hi, lo, C2)
    assert not torch.isfinite(C2).all()          # documents the failure mode the row maxima prevent


# ---------------------------------------------------------------- merged single-accumulator format (128x256 tiles)
def _split_rows(ops, W):
    hi, lo = torch.empty_like(W, dtype=torch.float16), torch.empty_like(W, dtype=torch.float16)
    inv = torch.empty(W.shape[0], device="cuda")
    ops.f16_split_rows(W, hi, lo, inv)
    return hi, lo, inv


def test_f16_split_rows_scales_each_row_into_fp16_range():
    from matinvent_b200 import ops
    W = _rand(300, 512, seed=40) * torch.logspace(-8, 6, 300).cuda()[:, None]
    W[7] = 0.0
    hi, lo, inv = _split_rows(ops, W)
    s = 1.0 / inv
    assert torch.equal(torch.log2(s), torch.log2(s).round())                    # powers of two
    top = (W * s[:, None]).abs().amax(dim=1)
    nz = W.abs().amax(dim=1) > 0
    assert bool(((top[nz] >= 2.0 ** 14) & (top[nz] < 2.0 ** 15)).all()) and float(s[7]) == 1.0
    rec = (hi.double() + lo.double()) * inv.double()[:, None]
    row_err = (rec - W.double()).abs().amax(dim=1) / W.abs().amax(dim=1).clamp_min(1e-300).double()
    assert float(row_err[nz].max()) <= 2.0 ** -21, float(row_err[nz].max())


@pytest.mark.parametrize("M,N,K", [(128, 256, 32), (128, 512, 768), (1, 8, 32), (300, 512, 512), (2643, 1024, 512),
                                   (257, 100, 512), (34445, 512, 768), (34445, 512, 512), (1000, 512, 1024)])
def test_tc_gemm_merged_fp32_grade(M, N, K):
    """merged format: three times as many truncating accumulations in the one accumulator -> looser bound (1e-5)"""
    from matinvent_b200 import ops
    A, W = _rand(M, K, seed=1), _rand(N, K, seed=2) * torch.logspace(-3, 3, N).cuda()[:, None]
    hi, lo, inv = _split_rows(ops, W)
    amax = A.abs().amax(dim=1).contiguous()
    C = torch.full((M, N), float("nan"), device="cuda")
    ops.tc_gemm(A, hi, lo, C, a_amax=amax, col_scale=inv, flags=ops.TC_MERGED)
    ref = A.double() @ W.double().t()
    # per output column (W rows differ by 1e6), relative to the largest |a|.|w| of the column: the forward error bound
    # of a dot product scales with sum |a_k w_k|, not with a possibly cancelled result
    col_err = (C.double() - ref).abs().amax(dim=0) / (A.double().abs() @ W.double().abs().t()).amax(dim=0)
    print("merged tc_gemm M=%d N=%d K=%d rel err %.2e" % (M, N, K, float(col_err.max())))
    assert float(col_err.max()) < 1e-5, float(col_err.max())


def test_tc_gemm_merged_needs_row_maxima():
    from matinvent_b200 import ops
    from matinvent_b200._lib import MatInventLibError
    A, W = _rand(64, 64, seed=1), _rand(256, 64, seed=2)
    hi, lo, inv = _split_rows(ops, W)
    with pytest.raises(MatInventLibError):
        ops.tc_gemm(A, hi, lo, torch.empty(64, 256, device="cuda"), col_scale=inv, flags=ops.TC_MERGED)


def test_tc_gemm_merged_presplit_fourier_edge_block():
    """the first per-edge block as CSPNet issues it in merged format: Phi from mi_edge_fourier scaled by 2^14 with an
    unscaled tail, per-row scaled weight, two gathers, pre-activation store, SiLU, row maxima"""
    from matinvent_b200 import ops
    Nn, F, H = 37, 128, 512
    g = torch.Generator().manual_seed(3)
    x = torch.rand(Nn, 3, generator=g).cuda()
    E = 5000
    src = torch.randint(0, Nn, (E,), generator=g).int().cuda()
    dst = torch.randint(0, Nn, (E,), generator=g).int().cuda()
    phi = torch.empty(E, 6 * F, device="cuda")
    phi_hi = torch.empty(E, 6 * F, device="cuda", dtype=torch.float16)
    phi_lo = torch.empty_like(phi_hi)
    ops.edge_fourier(x, src, dst, None, E, F, None, phi, phi_hi, phi_lo, op_scale=2.0 ** 14, lo_scale=1.0)
    assert float(((phi_hi.double() + phi_lo.double()) * 2.0 ** -14 - phi.double()).abs().max()) <= 2.0 ** -22
    W = _rand(H, 6 * F, seed=4) / (6 * F) ** 0.5
    hi, lo, inv = _split_rows(ops, W)
    PQ = _rand(Nn, 2 * H, seed=5)
    Z, C = torch.empty(E, H, device="cuda"), torch.empty(E, H, device="cuda")
    amax = torch.zeros(E, device="cuda")
    ops.tc_gemm_presplit(phi_hi, phi_lo, hi, lo, C, M=E, gathers=[(PQ[:, :H], src), (PQ[:, H:], dst)], z_out=Z,
                         act=ops.ACT_SILU, alpha=2.0 ** -14, col_scale=inv, amax_out=amax, flags=ops.TC_MERGED)
    z = phi.double() @ W.double().t() + PQ.double()[src.long(), :H] + PQ.double()[dst.long(), H:]
    assert rel_err(Z, z) < 5e-6
    assert rel_err(C, torch.nn.functional.silu(z)) < 5e-6
    assert torch.equal(amax, C.abs().amax(dim=1))


# ---------------------------------------------------------------- weight gradients: transposes + split-K accumulate
def test_transposes_for_weight_gradients():
    from matinvent_b200 import ops
    M, C = 1003, 200
    X = _rand(M, C, seed=50) * torch.logspace(-4, 4, C).cuda()[None, :]
    X[:, 5] = 0.0
    Mp = (M + 7) // 8 * 8
    XT = torch.zeros(C, Mp, device="cuda")
    ca = torch.zeros(C, device="cuda")
    ops.transpose_amax(X, M, XT=XT, col_amax=ca)
    assert torch.equal(XT[:, :M], X.t()) and float(XT[:, M:].abs().max()) == 0.0
    assert torch.equal(ca, X.abs().amax(dim=0))
    hi = torch.zeros(C, Mp, device="cuda", dtype=torch.float16)
    lo, inv = torch.zeros_like(hi), torch.zeros(C, device="cuda")
    ops.transpose_split(X, M, ca, hi, lo, inv)
    rec = (hi.double() + lo.double())[:, :M] * inv.double()[:, None]
    nz = ca > 0
    err = (rec - X.t().double()).abs().amax(dim=1) / ca.double().clamp_min(1e-300)
    assert float(err[nz].max()) <= 2.0 ** -21 and float(inv[5]) == 1.0


@pytest.mark.parametrize("M,N,K,ks", [(512, 512, 27860, 9), (512, 768, 9001, 4), (1024, 512, 8200, 18), (128, 256, 100, 2)])
def test_tc_gemm_split_k_accumulates_weight_gradient(M, N, K, ks):
    """dW += dY^T X as CSPNet._wgrad issues it: both operands transposed, K (the edge / node rows) cut into parts that
    are added to the gradient buffer"""
    from matinvent_b200 import ops
    dY = _rand(K, M, seed=60) * 1e-3
    X = _rand(K, N, seed=61) * torch.logspace(-2, 2, N).cuda()[None, :]
    G0 = _rand(M, N, seed=62) * 1e-6        # a previous accumulation, small enough not to dominate fp32 rounding
    G = G0.clone()
    Kp = (K + 7) // 8 * 8
    dyT = torch.zeros(M, Kp, device="cuda")
    ca, cx, inv = torch.zeros(M, device="cuda"), torch.zeros(N, device="cuda"), torch.zeros(N, device="cuda")
    hi = torch.zeros(N, Kp, device="cuda", dtype=torch.float16)
    lo = torch.zeros_like(hi)
    ops.transpose_amax(dY, K, XT=dyT, col_amax=ca)
    ops.transpose_amax(X, K, XT=None, col_amax=cx)
    ops.transpose_split(X, K, cx, hi, lo, inv)
    ops.tc_gemm(dyT, hi, lo, G, M=M, N=N, K=K, a_amax=ca, col_scale=inv, flags=ops.TC_MERGED, splitk=ks)
    ref = dY.double().t() @ X.double()
    col_err = ((G.double() - G0.double()) - ref).abs().amax(dim=0) / (dY.double().abs().t() @ X.double().abs()).amax(dim=0)
    print("split-K weight gradient M=%d N=%d K=%d ks=%d: %.2e" % (M, N, K, ks, float(col_err.max())))
    assert float(col_err.max()) < 1e-5


# ---------------------------------------------------------------- round 2: fused scatter-mean epilogue, LayerNorm -> pre-split operand
def _segments(sizes, seed=0):
    """consecutive-row segments: (idx [M] int32, weight [M] = 1 / segment length, ptr)"""
    idx = torch.repeat_interleave(torch.arange(len(sizes)), torch.tensor(sizes))
    w = (1.0 / torch.tensor(sizes, dtype=torch.float32))[idx]
    return idx.int().cuda(), w.cuda()


@pytest.mark.parametrize("sizes_kind,train", [("fc", False), ("fc", True), ("ragged", False), ("one", False)])
def test_tc_gemm_fused_scatter_mean(sizes_kind, train):
    """mi_tc_gemm with epi->scat_*: silu(A W^T + b) reduced to segment means in the epilogue (cspnet.py:73-79) against
    float64; segments straddle the 32-row windows and the 128-row tiles, rows past M are ignored, the reported maxima
    bound the means"""
    from matinvent_b200 import ops
    g = torch.Generator().manual_seed(3)
    if sizes_kind == "fc":          # n^2 edges per crystal in runs of n, like fc edges: 1..20 rows per segment
        ns = torch.randint(1, 21, (230,), generator=g).tolist()
        sizes = [n for n in ns for _ in range(n)]
    elif sizes_kind == "ragged":
        sizes = torch.randint(1, 60, (400,), generator=g).tolist()       # longer than a 32-row window too
    else:
        sizes = [1] * 517
    M, N, K = sum(sizes), 512, 512
    A = _rand(M, K, seed=4) * torch.exp(2 * _rand(M, 1, seed=5))
    W, bias = _rand(N, K, seed=6) / K ** 0.5, _rand(N, seed=7)
    mhi, mlo = torch.empty_like(W, dtype=torch.float16), torch.empty_like(W, dtype=torch.float16)
    inv = torch.empty(N, device="cuda")
    ops.f16_split_rows(W, mhi, mlo, inv)
    idx, w = _segments(sizes)
    S = len(sizes)
    out = torch.zeros(S, 2 * N, device="cuda")
    agg = out[:, N:]                                   # strided destination (cat[:, H:])
    amax = torch.zeros(S, device="cuda")
    Z = torch.empty(M, N, device="cuda") if train else None
    ops.tc_gemm(A, mhi, mlo, None, M=M, bias=bias, z_out=Z, act=ops.ACT_SILU, a_amax=A.abs().amax(1).contiguous(), col_scale=inv,
                flags=ops.TC_MERGED, scatter=(agg, idx, w, amax))
    z = A.double() @ W.double().t() + bias.double()
    y = torch.nn.functional.silu(z)
    ref = torch.zeros(S, N, dtype=torch.float64, device="cuda").index_add_(0, idx.long(), y)
    ref /= torch.tensor(sizes, dtype=torch.float64, device="cuda")[:, None]
    assert rel_err(agg, ref) < 6e-6, rel_err(agg, ref)
    assert float(out[:, :N].abs().max()) == 0.0                             # nothing written outside the destination
    if train:
        assert rel_err(Z, z) < 5e-6
    rowmax = ref.abs().amax(1).float()
    ymax = torch.zeros(S, dtype=torch.float64, device="cuda").index_reduce_(0, idx.long(), y.abs().amax(1), "amax").float()
    assert bool((amax >= rowmax * (1 - 1e-5)).all()) and bool((amax <= ymax * (1 + 1e-5)).all())
    # run to run bit-identical for fc-sized segments (at most two partial sums per destination element)
    if sizes_kind == "fc":
        out2 = torch.zeros(S, 2 * N, device="cuda")
        ops.tc_gemm(A, mhi, mlo, None, M=M, bias=bias, act=ops.ACT_SILU, a_amax=A.abs().amax(1).contiguous(), col_scale=inv,
                    flags=ops.TC_MERGED, scatter=(out2[:, N:], idx, w, None))
        assert torch.equal(out2[:, N:], agg)


def test_layernorm_split_feeds_presplit_gemm():
    """mi_layernorm_fwd_split: LayerNorm (cspnet.py:86-88) written as the pre-split fp16 operand of mi_tc_gemm_presplit
    with the row scale its a_amax implies; the GEMM on it against float64 LayerNorm @ W^T; rows spanning 1e-4..1e6"""
    from matinvent_b200 import ops
    rows, H, N = 777, 512, 1536
    x = _rand(rows, H, seed=1) * torch.exp(4 * _rand(rows, 1, seed=2)) * (1 + 0.3 * _rand(rows, 1, seed=3))
    gamma, beta = 1 + 0.3 * _rand(H, seed=4), 0.2 * _rand(H, seed=5)
    gamma[7] = 300.0                                   # one large affine weight: row maxima far above sqrt(H)
    W = _rand(N, H, seed=6) / H ** 0.5
    hi, lo = _split(ops, W)
    y = torch.empty(rows, 2 * H, device="cuda")
    yh, yl = torch.empty(rows, H, device="cuda", dtype=torch.float16), torch.empty(rows, H, device="cuda", dtype=torch.float16)
    amax = torch.empty(rows, device="cuda")
    tail = torch.full((rows, 2 * H), 7.0, device="cuda")
    mean, rstd = torch.empty(rows, device="cuda"), torch.empty(rows, device="cuda")
    ops.layernorm_fwd_split(x, gamma, beta, y[:, :H], yh, yl, amax, rows, H, zero_out=tail[:, H:], zero_cols=H, mean=mean, rstd=rstd)
    ref = torch.nn.functional.layer_norm(x.double(), (H,), gamma.double(), beta.double(), 1e-5)
    assert float(((y[:, :H].double() - ref).abs().amax(1) / ref.abs().amax(1)).max()) < 3e-6
    y2 = torch.empty(rows, H, device="cuda")
    ops.layernorm_fwd(x, gamma, beta, y2, rows, H)
    assert torch.equal(y2, y[:, :H])                                        # same arithmetic as the plain kernel
    assert torch.equal(amax, y2.abs().amax(1))
    assert float(tail[:, H:].abs().max()) == 0.0 and float(tail[:, :H].min()) == 7.0
    assert float(((mean.double() - x.double().mean(1)).abs() / x.double().abs().amax(1)).max()) < 1e-6
    # the operand pair reproduces the row to 2^-22 of its maximum once the row scale is undone
    e = torch.floor(torch.log2(amax)) - 14
    rec = (yh.double() + yl.double() / 2048) * torch.exp2(e.double())[:, None]
    assert float(((rec - y2.double()).abs() / amax.double()[:, None]).max()) < 2.0 ** -21
    Cb = _rand(9, N, seed=8)
    ib = torch.randint(0, 9, (rows,), generator=torch.Generator().manual_seed(9)).int().cuda()
    C = torch.empty(rows, N, device="cuda")
    ops.tc_gemm_presplit(yh, yl, hi, lo, C, M=rows, gathers=[(Cb, ib)], a_amax=amax)
    want = ref @ W.double().t() + Cb.double()[ib.long()]
    err = float(((C.double() - want).abs().amax(1) / want.abs().amax(1)).max())
    print("LN -> presplit GEMM: worst row error %.2e" % err)
    assert err < 6e-6, err


@pytest.mark.parametrize("num_atoms", [[4, 11, 20, 8], [20] * 13 + [1, 1, 3], "bench128", "bench190", "bench"])
def test_node_chain_matches_separate_kernels(gold_full, num_atoms):
    """mi_node_chain (one cluster launch per layer boundary: node_mlp.0 -> node_mlp.2 + residual -> next LayerNorm -> next
    P|Q|R GEMM) against the same forward through the separate kernels, and both against the oracle: row counts below,
    across and far above the 128-row blocks.  The library picks the transposed form (rows on the MMA's N side) while the row
    blocks are at most 64 rows — the first four cases: 8-, 16-, 48- and 64-row blocks — and the row-per-lane form above."""
    import numpy as np
    from oracle import diffcsp_oracle as O
    from test_gpu_parity import _full_module
    from matinvent_b200.models.diffcsp.sample import ATOM_DIST
    if isinstance(num_atoms, str):
        n = {"bench": 256, "bench128": 128, "bench190": 190}[num_atoms]
        num_atoms = np.random.RandomState(0).choice(21, n, p=ATOM_DIST["mp_20"]).tolist()
    na = torch.tensor(num_atoms)
    B, N = len(na), int(na.sum())
    g = torch.Generator().manual_seed(3)
    t = torch.randn(B, 256, generator=g)
    a = torch.randn(N, 100, generator=g)
    x = torch.rand(N, 3, generator=g)
    l = torch.randn(B, 3, 3, generator=g) + 4.0 * torch.eye(3)
    n2g = torch.repeat_interleave(torch.arange(B), na)
    m = _full_module(gold_full)
    outs = []
    for chain in (True, False):
        m.decoder.use_chain = chain
        with torch.no_grad():
            outs.append([o.clone() for o in m.decoder(t.cuda(), a.cuda(), x.cuda(), l.cuda(), na, n2g)])
    for u, v in zip(*outs):
        assert rel_err(u, v) < 3e-6, rel_err(u, v)
    sd = O.init_params(gold_full["hp"], gold_full["seeds"][0])
    with torch.no_grad():
        ref = O.cspnet_forward(sd, gold_full["hp"], t, a, x, l, na, n2g)
    errs = [rel_err(u, r) for u, r in zip(outs[0], ref)]
    print("node chain forward vs oracle (%d rows): %s" % (N, " ".join("%.2e" % e for e in errs)))
    assert max(errs) < 2e-5, errs
    # a second call on the same workspace gives the same bits (every buffer the chain reuses is re-initialised)
    m.decoder.use_chain = True
    with torch.no_grad():
        again = m.decoder(t.cuda(), a.cuda(), x.cuda(), l.cuda(), na, n2g)
    for u, v in zip(outs[0], again):
        assert torch.equal(u, v)


# ---------------------------------------------------------------- CTA-pair per-edge blocks (csrc/mi_edge.cu)
def test_edge_pair_blocks_tile_boundaries():
    """edge counts at and around the 256-row pair tile and the 128-row CTA half (1, 127, 128, 129, 255, 256, 257, 512 rows):
    rows past the last tile are clipped by the TMA stores and skipped by the scatter, nothing outside the outputs is touched"""
    import math
    from matinvent_b200 import ops
    H, K1, nn = 512, 768, 40
    g = torch.Generator().manual_seed(9)
    WF, W2 = _rand(H, K1, seed=6) / K1 ** 0.5, _rand(H, H, seed=7) / H ** 0.5
    fh, fl = torch.empty_like(WF, dtype=torch.float16), torch.empty_like(WF, dtype=torch.float16)
    wh, wl = torch.empty_like(W2, dtype=torch.float16), torch.empty_like(W2, dtype=torch.float16)
    finv, winv = torch.empty(H, device="cuda"), torch.empty(H, device="cuda")
    ops.f16_split_rows(WF, fh, fl, finv)
    ops.f16_split_rows(W2, wh, wl, winv)
    wfb = (math.sqrt(K1 / 2) * WF.norm(dim=1).max() * 1.001).reshape(1).contiguous()
    pqr = _rand(nn, 3 * H, seed=9).contiguous()
    amax_pq = pqr.abs().amax(1).contiguous()
    for E in (1, 127, 128, 129, 255, 256, 257, 512):
        src = torch.sort(torch.randint(0, nn, (E,), generator=g)).values
        dst = torch.randint(0, nn, (E,), generator=g)
        ang = (torch.rand(E, K1 // 2, generator=g) * 2 * math.pi).cuda()
        phi = torch.cat([ang.sin(), ang.cos()], dim=1).contiguous()
        ph = (phi * 2.0 ** 14).half()
        pl = (phi * 2.0 ** 14 - ph.float()).half()
        a1 = torch.full((2, E + 3, H), 7.0, device="cuda", dtype=torch.float16)          # three guard rows behind the tensor
        bound = torch.full((E + 3,), 7.0, device="cuda")
        ops.edge_block1(E, ph, pl, fh, fl, finv, 2.0 ** -14, pqr[:, :H], pqr[:, H:2 * H], src.int().cuda(), dst.int().cuda(), amax_pq,
                        wfb, a1[0][:E], a1[1][:E], bound[:E])
        assert float((a1[:, E:].float() - 7.0).abs().max()) == 0.0 and float((bound[E:] - 7.0).abs().max()) == 0.0
        z1 = phi.double() @ WF.double().t() + pqr[:, :H].double()[src.cuda()] + pqr[:, H:2 * H].double()[dst.cuda()]
        e8 = torch.floor(torch.log2(bound[:E])) - 14
        got1 = (a1[0][:E].double() + a1[1][:E].double()) * torch.exp2(e8.double())[:, None]
        err1 = float(((got1 - torch.nn.functional.silu(z1)).abs().amax(1) / z1.abs().amax(1)).max())
        cnt = torch.bincount(src, minlength=nn).clamp_min(1)
        w = (1.0 / cnt.float())[src].cuda()
        out = torch.zeros(nn + 2, H, device="cuda")
        ops.edge_block2(E, a1[0][:E], a1[1][:E], bound[:E], wh, wl, winv, None, out[:nn], src.int().cuda(), w, None)
        y = torch.nn.functional.silu(got1 @ W2.double().t())
        ref = torch.zeros(nn, H, dtype=torch.float64, device="cuda").index_add_(0, src.cuda(), y) / cnt.double().cuda()[:, None]
        err2 = float((out[:nn].double() - ref).abs().max() / ref.abs().max())
        assert err1 < 6e-6 and err2 < 8e-6 and float(out[nn:].abs().max()) == 0.0, (E, err1, err2)


@pytest.mark.parametrize("crystals,spread", [(3, 0.0), (40, 0.0), (230, 3.0), (900, 0.0)])
def test_edge_pair_blocks_vs_float64(crystals, spread):
    """mi_edge_block1 -> mi_edge_block2 (tcgen05.mma.cta_group::2, both operands staged by TMA, a1 handed over as an fp16
    pair scaled from an a-priori row bound) against float64 of cspnet.py:59-79: fc-like edge lists from one tile to many
    waves, rows past the last 256-row tile, node rows spanning e^(+-3 sigma) in magnitude"""
    import math
    from matinvent_b200 import ops
    g = torch.Generator().manual_seed(5)
    ns = torch.randint(1, 21, (crystals,), generator=g).tolist()
    nn = sum(ns)
    off = [0]
    for n in ns:
        off.append(off[-1] + n)
    src = torch.cat([torch.arange(off[b], off[b + 1]).repeat_interleave(ns[b]) for b in range(crystals)])
    dst = torch.cat([torch.arange(off[b], off[b + 1]).repeat(ns[b]) for b in range(crystals)])
    E, H, F = src.numel(), 512, 128
    K1 = 6 * F
    # Fourier-like operand: sin / cos pairs of unit norm, in mi_edge_fourier's merged format (scaled 2^14, unscaled tail)
    ang = (torch.rand(E, K1 // 2, generator=g) * 2 * math.pi).cuda()
    phi = torch.cat([ang.sin(), ang.cos()], dim=1).contiguous()
    ph = (phi * 2.0 ** 14).half()
    pl = (phi * 2.0 ** 14 - ph.float()).half()
    WF, W2, b2 = _rand(H, K1, seed=6) / K1 ** 0.5, _rand(H, H, seed=7) / H ** 0.5, 0.1 * _rand(H, seed=8)
    pqr = (_rand(nn, 3 * H, seed=9) * torch.exp(spread * _rand(nn, 1, seed=10))).contiguous()
    amax_pq = pqr.abs().amax(1).contiguous()

    def rows(W):
        hi, lo = torch.empty_like(W, dtype=torch.float16), torch.empty_like(W, dtype=torch.float16)
        inv = torch.empty(W.shape[0], device="cuda")
        ops.f16_split_rows(W, hi, lo, inv)
        return hi, lo, inv

    fh, fl, finv = rows(WF)
    wh, wl, winv = rows(W2)
    wfb = (math.sqrt(K1 / 2) * WF.norm(dim=1).max() * 1.001).reshape(1).contiguous()
    a1 = torch.empty(2, E, H, device="cuda", dtype=torch.float16)
    bound = torch.full((E,), float("nan"), device="cuda")
    srci, dsti = src.int().cuda(), dst.int().cuda()
    ops.edge_block1(E, ph, pl, fh, fl, finv, 2.0 ** -14, pqr[:, :H], pqr[:, H:2 * H], srci, dsti, amax_pq, wfb, a1[0], a1[1], bound)
    z1 = phi.double() @ WF.double().t() + pqr[:, :H].double()[src.cuda()] + pqr[:, H:2 * H].double()[dst.cuda()]
    r1 = torch.nn.functional.silu(z1)
    assert bool((bound >= r1.abs().amax(1).float()).all())                    # it is a bound
    e8 = torch.floor(torch.log2(bound)) - 14
    got1 = (a1[0].double() + a1[1].double()) * torch.exp2(e8.double())[:, None]
    err1 = float(((got1 - r1).abs().amax(1) / z1.abs().amax(1)).max())
    slack = float((bound / r1.abs().amax(1).float().clamp_min(1e-30)).median())
    print("edge block 1: E=%d worst row error %.2e (median bound / max = %.1f)" % (E, err1, slack))
    assert err1 < 6e-6, err1
    # block 2 on block 1's own output: means over the source node's edges
    seg = torch.repeat_interleave(torch.arange(nn), torch.tensor([n for n in ns for _ in range(n)]))
    w = (1.0 / torch.tensor([n for n in ns for _ in range(n)], dtype=torch.float32))[seg].cuda()
    out = torch.zeros(nn, 2 * H, device="cuda")
    agg, amax = out[:, H:], torch.zeros(nn, device="cuda")
    ops.edge_block2(E, a1[0], a1[1], bound, wh, wl, winv, b2, agg, srci, w, amax)
    y = torch.nn.functional.silu(got1 @ W2.double().t() + b2.double())
    ref = torch.zeros(nn, H, dtype=torch.float64, device="cuda").index_add_(0, seg.cuda(), y) * \
        (1.0 / torch.tensor(ns, dtype=torch.float64).repeat_interleave(torch.tensor(ns))).cuda()[:, None]
    err2 = float(((agg.double() - ref).abs().amax(1) / ref.abs().amax(1)).max())
    print("edge block 2: worst row error %.2e" % err2)
    assert err2 < 8e-6, err2
    assert float(out[:, :H].abs().max()) == 0.0
    assert bool((amax >= ref.abs().amax(1).float() * (1 - 1e-5)).all())
    # against the single-CTA kernels on the same operands (same format, same products): rounding-level agreement
    a_ref = torch.empty(E, H, device="cuda")
    am = torch.zeros(E, device="cuda")
    ops.tc_gemm_presplit(ph, pl, fh, fl, a_ref, M=E, alpha=2.0 ** -14, col_scale=finv, flags=ops.TC_MERGED,
                         gathers=[(pqr[:, :H], srci), (pqr[:, H:2 * H], dsti)], act=ops.ACT_SILU, amax_out=am)
    assert float(((got1 - a_ref.double()).abs().amax(1) / z1.abs().amax(1)).max()) < 4e-6
