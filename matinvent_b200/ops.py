"""Tensor-level wrappers over the C ABI (include/matinvent_b200.h).  torch is used for device memory
and streams only; every arithmetic op below runs in libmatinvent_b200.so."""
import ctypes as C

import torch

from . import _lib
from ._lib import Epilogue, check

ACT_NONE, ACT_SILU, ACT_DSILU = 0, 1, 2


def lib():
    return _lib.load()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _p(t):
    return None if t is None else t.data_ptr()


def _f32(t, name="tensor"):
    if t is None:
        return
    if t.dtype != torch.float32 or not t.is_cuda:
        raise TypeError("%s must be a CUDA float32 tensor (got %s on %s)" % (name, t.dtype, t.device))


def _i32(t, name="tensor"):
    if t is None:
        return
    if t.dtype != torch.int32 or not t.is_cuda:
        raise TypeError("%s must be a CUDA int32 tensor (got %s on %s)" % (name, t.dtype, t.device))


def _ld(t):
    """leading dimension of a 2-D row-major view (last dim contiguous)."""
    if t.dim() == 1:
        return t.shape[0]
    if t.stride(-1) != 1 and t.shape[-1] != 1:
        raise ValueError("last dimension must be contiguous")
    return t.stride(0) if t.shape[0] > 1 else max(t.shape[-1], t.stride(0))


def sgemm(A, B, C_, transA=False, transB=True, M=None, N=None, K=None, bias=None, gathers=(), z_out=None,
          z_in=None, resid=None, act=ACT_NONE, alpha=1.0, beta=0.0, splitk=1, amax_out=None, a_amax=None):
    """C = epilogue(alpha * op(A) @ op(B)); see mi_sgemm.  A, B, C are 2-D row-major views.
    gathers: up to three (tensor, index-or-None) pairs added row-wise.  a_amax is ignored (fp32 has the range)."""
    for t in (A, B, C_, bias, z_out, z_in, resid):
        _f32(t)
    if M is None:
        M = A.shape[1] if transA else A.shape[0]
    if K is None:
        K = A.shape[0] if transA else A.shape[1]
    if N is None:
        N = B.shape[0] if transB else B.shape[1]
    e = _epilogue(bias, gathers, z_out, z_in, resid, act, alpha, beta, splitk, amax_out, None)
    check(lib().mi_sgemm(int(transA), int(transB), M, N, K, A.data_ptr(), _ld(A), B.data_ptr(), _ld(B),
                         C_.data_ptr(), _ld(C_), C.byref(e), _stream()), "mi_sgemm")
    return C_


def _epilogue(bias, gathers, z_out, z_in, resid, act, alpha, beta, splitk, amax_out=None, a_amax=None, col_scale=None,
              scatter=None):
    e = Epilogue()
    if scatter is not None:          # (out [S, >= N], idx [M] int32, weight [M], amax [S] or None)
        so, si, sw, sa = scatter
        _f32(so), _i32(si), _f32(sw), _f32(sa)
        e.scat_out, e.scat_ld, e.scat_idx, e.scat_w, e.scat_amax = _p(so), _ld(so), _p(si), _p(sw), _p(sa)
    _f32(amax_out), _f32(a_amax), _f32(col_scale)
    e.amax_out, e.a_amax, e.col_scale = _p(amax_out), _p(a_amax), _p(col_scale)
    e.bias = _p(bias)
    g = list(gathers) + [(None, None)] * (3 - len(gathers))
    for k, (src, idx) in enumerate(g[:3]):
        _f32(src)
        _i32(idx)
        setattr(e, "g%d" % (k + 1), _p(src))
        setattr(e, "g%d_idx" % (k + 1), _p(idx))
        setattr(e, "g%d_ld" % (k + 1), _ld(src) if src is not None else 0)
    e.z_out, e.z_ld = _p(z_out), (_ld(z_out) if z_out is not None else 0)
    e.z_in, e.zin_ld = _p(z_in), (_ld(z_in) if z_in is not None else 0)
    e.resid, e.resid_ld = _p(resid), (_ld(resid) if resid is not None else 0)
    e.act, e.alpha, e.beta, e.splitk = act, alpha, beta, splitk
    return e


_SM_COUNT = None


def sm_count():
    """SMs of the current device (mi_device_info), cached"""
    global _SM_COUNT
    if _SM_COUNT is None:
        sm, ma, mi = C.c_int(0), C.c_int(0), C.c_int(0)
        rc = lib().mi_device_info(C.byref(sm), C.byref(ma), C.byref(mi))
        if rc != 0:
            raise RuntimeError("mi_device_info failed (rc=%d)" % rc)
        _SM_COUNT = int(sm.value)
    return _SM_COUNT


TC_MERGED = 1      # MI_TC_MERGED: single-accumulator 128x256 tiles, operands (s x) = hi + lo with unscaled lo


def f16_split(w, hi, lo, scale=1.0, lo_scale=2048.0):
    """hi = fp16(scale w), lo = fp16((scale w - hi) * lo_scale): operands of the split-precision tensor-core GEMM
    ((1, 2^11): separate-accumulator format; (s, 1): merged format, s a power of two with max |s w| in [2^14, 2^15))."""
    _f32(w)
    if hi.dtype != torch.float16 or lo.dtype != torch.float16:
        raise TypeError("hi/lo must be float16")
    check(lib().mi_f16_split(_p(w), _p(hi), _p(lo), w.numel(), scale, lo_scale, _stream()), "mi_f16_split")


def f16_split_rows(w, hi, lo, inv_scale):
    """merged-format split of a weight matrix [rows, cols] with one power-of-two scale per row, chosen on the device;
    inv_scale [rows] is what tc_gemm's col_scale takes.  See mi_f16_split_rows."""
    _f32(w), _f32(inv_scale)
    if hi.dtype != torch.float16 or lo.dtype != torch.float16 or _ld(hi) != _ld(w) or _ld(lo) != _ld(w):
        raise TypeError("hi/lo must be float16 with the leading dimension of w")
    check(lib().mi_f16_split_rows(_p(w), w.shape[0], w.shape[1], _ld(w), _p(hi), _p(lo), _p(inv_scale), _stream()),
          "mi_f16_split_rows")


def transpose_amax(X, M, XT=None, col_amax=None):
    """XT[c][m] = X[m][c] for m < M (XT [C, >= M] fp32, nullable) and col_amax[c] = max(col_amax[c], max_m |X[m][c]|)"""
    _f32(X), _f32(XT), _f32(col_amax)
    check(lib().mi_transpose_amax(_p(X), _ld(X), M, X.shape[1], _p(XT), _ld(XT) if XT is not None else 0, _p(col_amax),
                                  _stream()), "mi_transpose_amax")


def transpose_split(X, M, col_amax, hi, lo, inv_scale):
    """hi/lo[c][m] = merged-format fp16 split of X[m][c] scaled per column from col_amax; inv_scale[c] = 1 / scale"""
    _f32(X), _f32(col_amax), _f32(inv_scale)
    if hi.dtype != torch.float16 or lo.dtype != torch.float16 or _ld(hi) != _ld(lo):
        raise TypeError("hi/lo must be float16 with equal leading dimensions")
    check(lib().mi_transpose_split(_p(X), _ld(X), M, X.shape[1], _p(col_amax), _p(hi), _p(lo), _ld(hi), _p(inv_scale),
                                   _stream()), "mi_transpose_split")


def merged_scale(w):
    """the power of two s with max |s w| in [2^14, 2^15) (1.0 for an all-zero tensor); one host sync"""
    m = float(w.abs().max())
    if m == 0.0 or m != m or m == float("inf"):
        return 1.0
    import math
    return 2.0 ** (14 - math.frexp(m)[1] + 1)


def tc_ok(A, W):
    """operands usable by the TMA / tcgen05 path: 16-byte aligned rows (fp32 A: ld % 4, fp16 W copies: ld % 8)"""
    wa = 16 if W.dtype == torch.float16 else 32      # fp16 copy: its own alignment; fp32 master: the copy's is half
    return (_ld(A) % 4 == 0 and _ld(W) % 8 == 0 and A.data_ptr() % 16 == 0 and W.data_ptr() % wa == 0)


def tc_gemm(A, W_hi, W_lo, C_, M=None, N=None, K=None, bias=None, gathers=(), z_out=None, z_in=None, resid=None,
            act=ACT_NONE, alpha=1.0, beta=0.0, amax_out=None, a_amax=None, flags=0, col_scale=None, splitk=1, scatter=None):
    """C = epilogue(alpha * A @ W^T) on the tensor cores (split FP16); W_hi/W_lo from f16_split.  See mi_tc_gemm.
    splitk > 1: C += alpha * A @ W^T with K cut into splitk parts (plain epilogue only).
    scatter = (out, idx, weight, amax): rows are reduced by segment into `out` instead of being stored (C_ may be None)."""
    for t in (A, C_, bias, z_out, z_in, resid):
        _f32(t)
    M = A.shape[0] if M is None else M
    K = A.shape[1] if K is None else K
    N = W_hi.shape[0] if N is None else N
    e = _epilogue(bias, gathers, z_out, z_in, resid, act, alpha, beta, splitk, amax_out, a_amax, col_scale, scatter)
    check(lib().mi_tc_gemm(M, N, K, A.data_ptr(), _ld(A), W_hi.data_ptr(), W_lo.data_ptr(), _ld(W_hi), _p(C_),
                           _ld(C_) if C_ is not None else N, C.byref(e), flags, _stream()), "mi_tc_gemm")
    return C_


def tc_gemm_presplit(A_hi, A_lo, W_hi, W_lo, C_, M=None, N=None, K=None, bias=None, gathers=(), z_out=None, resid=None,
                     act=ACT_NONE, alpha=1.0, amax_out=None, flags=0, col_scale=None, a_amax=None):
    """tc_gemm with A given as fp16 (hi, scaled lo) arrays from its producer.  See mi_tc_gemm_presplit.
    a_amax: the row maxima the producer scaled the rows by (layernorm_fwd_split): result rows are scaled back."""
    for t in (C_, bias, z_out, resid):
        _f32(t)
    M = A_hi.shape[0] if M is None else M
    K = A_hi.shape[1] if K is None else K
    N = W_hi.shape[0] if N is None else N
    e = _epilogue(bias, gathers, z_out, None, resid, act, alpha, 0.0, 1, amax_out, a_amax, col_scale)
    check(lib().mi_tc_gemm_presplit(M, N, K, A_hi.data_ptr(), A_lo.data_ptr(), _ld(A_hi), W_hi.data_ptr(), W_lo.data_ptr(),
                                    _ld(W_hi), C_.data_ptr(), _ld(C_), C.byref(e), flags, _stream()), "mi_tc_gemm_presplit")
    return C_


def fc_edges(node_off, edge_off, B, N, E, edge_src, edge_dst, edge_graph, seg_ptr, dst_ptr, dst_perm, node_graph):
    for t in (node_off, edge_off, edge_src, edge_dst, edge_graph, seg_ptr, dst_ptr, dst_perm, node_graph):
        _i32(t)
    check(lib().mi_fc_edges(_p(node_off), _p(edge_off), B, N, E, _p(edge_src), _p(edge_dst), _p(edge_graph),
                            _p(seg_ptr), _p(dst_ptr), _p(dst_perm), _p(node_graph), _stream()), "mi_fc_edges")


def edge_fourier(x, edge_src, edge_dst, cell_off, E, F, frac_diff, phi, phi_hi=None, phi_lo=None, op_scale=1.0,
                 lo_scale=2048.0):
    _f32(x), _f32(phi), _f32(cell_off), _f32(frac_diff), _i32(edge_src), _i32(edge_dst)
    ld = _ld(phi) if phi is not None else _ld(phi_hi)
    check(lib().mi_edge_fourier(_p(x), _p(edge_src), _p(edge_dst), _p(cell_off), E, F, _p(frac_diff), _p(phi),
                                ld, _p(phi_hi), _p(phi_lo), op_scale, lo_scale, _stream()), "mi_edge_fourier")


def segment_reduce(X, ptr, out, S, H, perm=None, mean=True, accumulate=False, amax_out=None):
    _f32(X), _f32(out), _i32(ptr), _i32(perm), _f32(amax_out)
    check(lib().mi_segment_reduce(_p(X), _ld(X), _p(ptr), _p(perm), _p(out), _ld(out), S, H, int(mean),
                                  int(accumulate), _p(amax_out), _stream()), "mi_segment_reduce")
    return out


def gather_rows_dsilu(dOut, idx, ptr, z, dX, E, H, amax_out=None):
    _f32(dOut), _f32(z), _f32(dX), _i32(idx), _i32(ptr), _f32(amax_out)
    check(lib().mi_gather_rows_dsilu(_p(dOut), _ld(dOut), _p(idx), _p(ptr), _p(z), _ld(z) if z is not None else 0,
                                     _p(dX), _ld(dX), E, H, _p(amax_out), _stream()), "mi_gather_rows_dsilu")
    return dX


def row_amax(X, rows, cols, out):
    _f32(X), _f32(out)
    check(lib().mi_row_amax(_p(X), _ld(X), rows, cols, _p(out), _stream()), "mi_row_amax")
    return out


def colsum(X, M, N, out, accumulate=True):
    _f32(X), _f32(out)
    check(lib().mi_colsum(_p(X), _ld(X), M, N, _p(out), int(accumulate), _stream()), "mi_colsum")
    return out


def layernorm_fwd(x, gamma, beta, y, rows, H, mean=None, rstd=None, eps=1e-5, amax_out=None):
    for t in (x, gamma, beta, y, mean, rstd, amax_out):
        _f32(t)
    check(lib().mi_layernorm_fwd(_p(x), _ld(x), _p(gamma), _p(beta), _p(y), _ld(y), _p(mean), _p(rstd), rows, H,
                                 eps, _p(amax_out), _stream()), "mi_layernorm_fwd")
    return y


def layernorm_fwd_split(x, gamma, beta, y, y_hi, y_lo, amax, rows, H, zero_out=None, zero_cols=0, mean=None, rstd=None, eps=1e-5):
    """LayerNorm written as the pre-split fp16 operand pair of tc_gemm_presplit (+ row maxima for its a_amax); y (fp32)
    optional; zero_out: a companion row block zeroed in the same pass.  See mi_layernorm_fwd_split."""
    for t in (x, gamma, beta, y, amax, zero_out, mean, rstd):
        _f32(t)
    if y_hi.dtype != torch.float16 or y_lo.dtype != torch.float16 or _ld(y_hi) != _ld(y_lo):
        raise TypeError("y_hi / y_lo must be float16 with equal leading dimensions")
    check(lib().mi_layernorm_fwd_split(_p(x), _ld(x), _p(gamma), _p(beta), _p(y), _ld(y) if y is not None else 0, _p(y_hi),
                                       _p(y_lo), _ld(y_hi), _p(amax), _p(zero_out), _ld(zero_out) if zero_out is not None else 0,
                                       zero_cols, _p(mean), _p(rstd), rows, H, eps, _stream()), "mi_layernorm_fwd_split")


def layernorm_bwd(dy, x, gamma, mean, rstd, dx, dgamma, dbeta, rows, H, accumulate_dx=False):
    for t in (dy, x, gamma, mean, rstd, dx, dgamma, dbeta):
        _f32(t)
    check(lib().mi_layernorm_bwd(_p(dy), _ld(dy), _p(x), _ld(x), _p(gamma), _p(mean), _p(rstd), _p(dx),
                                 _ld(dx) if dx is not None else 0, int(accumulate_dx), _p(dgamma), _p(dbeta),
                                 rows, H, _stream()), "mi_layernorm_bwd")


def lattice_ip(L, ips, B):
    _f32(L), _f32(ips)
    check(lib().mi_lattice_ip(_p(L), _p(ips), B, _stream()), "mi_lattice_ip")
    return ips


def lattice_linear(L, W, bias, out, B, H, n_sets=1, w_stride=0, bias_stride=0, out_stride=0):
    """out[s][b] = bias_s + vec(L_b L_b^T) W_s^T for n_sets weight sets spaced by the given element strides
    (W, bias: views of set 0; out: view of set 0 with row stride _ld(out)).  See mi_lattice_linear."""
    for t in (L, W, bias, out):
        _f32(t)
    check(lib().mi_lattice_linear(_p(L), _p(W), _p(bias), _p(out), _ld(out), B, H, n_sets, w_stride, bias_stride,
                                  out_stride, _stream()), "mi_lattice_linear")
    return out


def bmm3(A, L, out, B, transL=False):
    _f32(A), _f32(L), _f32(out)
    check(lib().mi_bmm3(_p(A), _p(L), _p(out), B, int(transL), _stream()), "mi_bmm3")
    return out


def time_embed(t, freq, B, dim, out):
    _i32(t), _f32(freq), _f32(out)
    check(lib().mi_time_embed(_p(t), _p(freq), B, dim, _p(out), _stream()), "mi_time_embed")
    return out


def lattice_params_to_matrix(lengths, angles, L, B):
    _f32(lengths), _f32(angles), _f32(L)
    check(lib().mi_lattice_params_to_matrix(_p(lengths), _p(angles), _p(L), B, _stream()), "mi_lattice_params_to_matrix")
    return L


def lattice_matrix_to_params(L, lengths, angles, B):
    _f32(lengths), _f32(angles), _f32(L)
    check(lib().mi_lattice_matrix_to_params(_p(L), _p(lengths), _p(angles), B, _stream()), "mi_lattice_matrix_to_params")


def argmax_rows(a, rows, cols, out, add=0):
    _f32(a), _i32(out)
    check(lib().mi_argmax_rows(_p(a), _ld(a), rows, cols, add, _p(out), _stream()), "mi_argmax_rows")
    return out


def reverse_corrector(x, pred_x, z_x, x_half, N, coef, t_dev=None, t_host=0):
    for t in (x, pred_x, z_x, x_half, coef):
        _f32(t)
    _i32(t_dev)
    check(lib().mi_reverse_corrector(_p(x), _p(pred_x), _p(z_x), _p(x_half), N, _p(coef), _p(t_dev), t_host,
                                     _stream()), "mi_reverse_corrector")


def reverse_predictor(x_half, pred_x, z_x, x, N, l, pred_l, z_l, B, a, pred_a, z_a, A, coef, t_dev=None, t_host=0):
    for t in (x_half, pred_x, z_x, x, l, pred_l, z_l, a, pred_a, z_a, coef):
        _f32(t)
    _i32(t_dev)
    check(lib().mi_reverse_predictor(_p(x_half), _p(pred_x), _p(z_x), _p(x), N, _p(l), _p(pred_l), _p(z_l), B, _p(a),
                                     _p(pred_a), _p(z_a), A, _p(coef), _p(t_dev), t_host, _stream()),
          "mi_reverse_predictor")


def sampler_step_begin(t_dev, ttab, temb, B, T):
    _i32(t_dev), _f32(ttab), _f32(temb)
    check(lib().mi_sampler_step_begin(_p(t_dev), _p(ttab), _p(temb), B, T, _stream()), "mi_sampler_step_begin")


def sampler_step_end(t_dev):
    _i32(t_dev)
    check(lib().mi_sampler_step_end(_p(t_dev), _stream()), "mi_sampler_step_end")


def add_noise(L0, x0, Z, z_l, z_x, z_a, B, N, A, coef, l_t, x_t, a_t, tar_x, t_dev=None, t_host=0):
    for t in (L0, x0, z_l, z_x, z_a, l_t, x_t, a_t, tar_x, coef):
        _f32(t)
    _i32(Z), _i32(t_dev)
    check(lib().mi_add_noise(_p(L0), _p(x0), _p(Z), _p(z_l), _p(z_x), _p(z_a), B, N, A, _p(coef), _p(t_dev), t_host,
                             _p(l_t), _p(x_t), _p(a_t), _p(tar_x), _stream()), "mi_add_noise")


def rl_loss(pred, tgt, prior, node_off, B, A, costs, w_loss, w_kl, scale, loss, kl, grads, stats=None):
    """pred/tgt/prior/grads: triples (l, x, a) or None."""
    tl, tx, ta = tgt if tgt is not None else (None, None, None)
    ql, qx, qa = prior if prior is not None else (None, None, None)
    dl, dx, da = grads if grads is not None else (None, None, None)
    for t in list(pred) + [tl, tx, ta, ql, qx, qa, dl, dx, da, w_loss, w_kl, loss, kl, stats]:
        _f32(t)
    _i32(node_off)
    check(lib().mi_rl_loss(_p(pred[0]), _p(pred[1]), _p(pred[2]), _p(tl), _p(tx), _p(ta), _p(ql), _p(qx), _p(qa),
                           _p(node_off), B, A, costs[0], costs[1], costs[2], _p(w_loss), _p(w_kl), scale, _p(loss),
                           _p(kl), _p(dl), _p(dx), _p(da), _p(stats), _stream()), "mi_rl_loss")


def adam_step(p, g, m, v, lr, step, b1=0.9, b2=0.999, eps=1e-8, grad_scale=1.0, zero_grad=True):
    for t in (p, g, m, v):
        _f32(t)
    check(lib().mi_adam_step(_p(p), _p(g), _p(m), _p(v), p.numel(), lr, b1, b2, eps, step, grad_scale,
                             int(zero_grad), _stream()), "mi_adam_step")


def philox_normal(out, seed, offset=0, offset_dev=None, advance=False):
    _f32(out)
    check(lib().mi_philox_normal(_p(out), out.numel(), seed, offset, _p(offset_dev), int(advance), _stream()),
          "mi_philox_normal")
    return out


def philox_uniform(out, seed, offset=0, offset_dev=None, advance=False):
    _f32(out)
    check(lib().mi_philox_uniform(_p(out), out.numel(), seed, offset, _p(offset_dev), int(advance), _stream()),
          "mi_philox_uniform")
    return out


def radius_graph_pbc(x, L, node_off, B, N, max_n, max_neighbors, cap, edge_dst, cell_off, deg, overflow):
    _f32(x), _f32(L), _f32(cell_off), _i32(node_off), _i32(edge_dst), _i32(deg), _i32(overflow)
    check(lib().mi_radius_graph_pbc(_p(x), _p(L), _p(node_off), B, N, max_n, max_neighbors, cap, _p(edge_dst), _p(cell_off),
                                    _p(deg), _p(overflow), _stream()), "mi_radius_graph_pbc")


def compact_edges(deg, N, cap, edge_dst_pad, cell_pad, node_graph, seg_ptr, edge_src, edge_dst, edge_graph, cell_off,
                  E_cap):
    check(lib().mi_compact_edges(_p(deg), N, cap, _p(edge_dst_pad), _p(cell_pad), _p(node_graph), _p(seg_ptr),
                                 _p(edge_src), _p(edge_dst), _p(edge_graph), _p(cell_off), E_cap, _stream()),
          "mi_compact_edges")


def build_dst_csr(seg_ptr, edge_dst, N, E_cap, dst_ptr, dst_perm, work):
    check(lib().mi_build_dst_csr(_p(seg_ptr), _p(edge_dst), N, E_cap, _p(dst_ptr), _p(dst_perm), _p(work), _stream()),
          "mi_build_dst_csr")


def replay_select(keys, rewards, n, buffer_size, cutoff, out_idx, out_count):
    check(lib().mi_replay_select(_p(keys), _p(rewards), n, buffer_size, cutoff, _p(out_idx), _p(out_count), _stream()),
          "mi_replay_select")


def composition_key(Z, node_off, B, keys):
    check(lib().mi_composition_key(_p(Z), _p(node_off), B, _p(keys), _stream()), "mi_composition_key")


def validity_prefilter(frac, L, lengths, node_off, B, mask, dmin=None, max_len=25.0, min_dist=0.5, min_vol=0.1, hard_len=40.0):
    _f32(frac), _f32(L), _f32(lengths), _f32(dmin), _i32(node_off), _i32(mask)
    check(lib().mi_validity_prefilter(_p(frac), _p(L), _p(lengths), _p(node_off), B, max_len, min_dist, min_vol, hard_len,
                                      _p(mask), _p(dmin), _stream()), "mi_validity_prefilter")
    return mask


def composition_reward(Z, node_off, B, tables, mass, modes, targets, minv, maxv, tval, weight, reduce, props, rewards, failed):
    _i32(Z), _i32(node_off), _i32(failed)
    for t in (tables, mass, props, rewards):
        if t.dtype != torch.float64 or not t.is_cuda:
            raise TypeError("tables / mass / props / rewards must be CUDA float64 tensors")
    P = len(modes)
    ia, da = (C.c_int * P), (C.c_double * P)
    check(lib().mi_composition_reward(_p(Z), _p(node_off), B, _p(tables), _p(mass), P, ia(*modes), ia(*targets), da(*minv),
                                      da(*maxv), da(*tval), da(*weight), reduce, _p(props), _p(rewards), _p(failed), _stream()),
          "mi_composition_reward")


def node_chain(M, H, agg, amax_agg, zero_agg, xs, ys, wb_hi, wb_lo, bn1, R, amax_pqr, bounds, w2_hi, w2_lo, bn2, h_in, h, ln=None):
    """node_mlp.0 -> node_mlp.2 + residual (h = h_in + ...; h may be h_in) -> [next layer's LayerNorm + P|Q|R GEMM] in one launch.
    xs, ys = (hi, lo) fp16 [M, H] scratch operand pairs; bounds = 3 device floats (see mi_node_chain);
    ln = (gamma, beta, eps, wpqr_hi, wpqr_lo, cb [B, 3H], node_graph, pqr [M, 3H], amax_next [M]) or None (last layer)."""
    for t in (agg, amax_agg, bn1, R, amax_pqr, bounds, bn2, h_in, h):
        _f32(t)
    for t in (*xs, *ys):
        if t.dtype != torch.float16 or not t.is_contiguous() or t.shape[1] != H:
            raise TypeError("xs / ys must be contiguous fp16 [M, H]")
    if ln is not None:
        g, b, eps, ph, pl, cb, ng, pqr, amax_next = ln
        _f32(g), _f32(b), _f32(cb), _f32(pqr), _i32(ng), _f32(amax_next)
        tail = (_p(g), _p(b), float(eps), _p(ph), _p(pl), _p(cb), _ld(cb), _p(ng), _p(pqr), _ld(pqr), _p(amax_next))
    else:
        tail = (None, None, 1e-5, None, None, None, 0, None, None, 0, None)
    check(lib().mi_node_chain(M, H, 3 if ln is not None else 2, _p(agg), _ld(agg), _p(amax_agg), int(bool(zero_agg)),
                              _p(xs[0]), _p(xs[1]), _p(ys[0]), _p(ys[1]), _p(wb_hi), _p(wb_lo), _ld(wb_hi), _p(bn1), _p(R), _ld(R),
                              _p(amax_pqr), _p(bounds), _p(w2_hi), _p(w2_lo), _p(bn2), _p(h_in), _ld(h_in), _p(h), _ld(h),
                              *tail, _stream()), "mi_node_chain")


def edge_block1(E, phi_hi, phi_lo, w_hi, w_lo, col_scale, alpha, P, Q, src, dst, amax_pq, wf_bound, a_hi, a_lo, a_bound):
    """first per-edge block on CTA pairs: (a_hi, a_lo) = split(silu(alpha phi W^T col_scale + P[src] + Q[dst])).  See mi_edge_block1."""
    _f32(col_scale), _f32(P), _f32(Q), _f32(amax_pq), _f32(wf_bound), _f32(a_bound), _i32(src), _i32(dst)
    N, K = w_hi.shape
    check(lib().mi_edge_block1(E, N, K, _p(phi_hi), _p(phi_lo), _ld(phi_hi), _p(w_hi), _p(w_lo), _ld(w_hi), _p(col_scale),
                               float(alpha), _p(P), _p(Q), _ld(P), _p(src), _p(dst), _p(amax_pq), _p(wf_bound), _p(a_hi), _p(a_lo),
                               _ld(a_hi), _p(a_bound), _stream()), "mi_edge_block1")


def edge_block2(E, a_hi, a_lo, a_bound, w_hi, w_lo, col_scale, bias, scat_out, scat_idx, scat_w, scat_amax):
    """second per-edge block + scatter-mean on CTA pairs.  See mi_edge_block2."""
    _f32(a_bound), _f32(col_scale), _f32(bias), _f32(scat_out), _i32(scat_idx), _f32(scat_w), _f32(scat_amax)
    N, K = w_hi.shape
    check(lib().mi_edge_block2(E, N, K, _p(a_hi), _p(a_lo), _ld(a_hi), _p(a_bound), _p(w_hi), _p(w_lo), _ld(w_hi), _p(col_scale),
                               _p(bias), _p(scat_out), _ld(scat_out), _p(scat_idx), _p(scat_w), _p(scat_amax), _stream()),
          "mi_edge_block2")


def output_heads(h, node_off, B, H, ln_g, ln_b, coord_w, pred_x, type_w, type_b, pred_a, lattice_w, L, ip, pred_l, eps=1e-5):
    """final LayerNorm + coordinate / type / lattice heads in one launch (any output may be None).  See mi_output_heads."""
    for t in (h, ln_g, ln_b, coord_w, pred_x, type_w, type_b, pred_a, lattice_w, L, pred_l):
        _f32(t)
    _i32(node_off)
    A = pred_a.shape[1] if pred_a is not None else 0
    check(lib().mi_output_heads(_p(h), _ld(h), _p(node_off), B, H, _p(ln_g), _p(ln_b), float(eps), _p(coord_w), _p(pred_x),
                                _p(type_w), _p(type_b), A, _p(pred_a), _p(lattice_w), _p(L), int(bool(ip)), _p(pred_l), _stream()),
          "mi_output_heads")


def weighted_field_sum(fields, weights, out):
    """out[b] = sum_k weights[k] * fields[k][b]  (<= 4 per-sample loss vectors).  See mi_weighted_field_sum."""
    for t in list(fields) + [out]:
        _f32(t)
    F = len(fields)
    ptrs = [_p(t) for t in fields] + [None] * (4 - F)
    check(lib().mi_weighted_field_sum(out.numel(), F, *ptrs, (C.c_float * F)(*[float(w) for w in weights]), _p(out), _stream()),
          "mi_weighted_field_sum")
    return out
