/* libmatinvent_b200 — C ABI of the B200-native DiffCSP/MatInvent hot path.
 *
 * The reference (schwallergroup/matinvent) is pure Python and has NO FFI of its own: its boundary is
 * the duck-typed plugin API of models/suite/base.py:30-59 and pipeline/base.py:26-142.  This header is
 * the operator ABI that our host-side mirror of that plugin API (matinvent_b200/models, /pipeline,
 * /memory) binds through ctypes; every entry point names the reference code it replaces.
 *
 * Conventions
 *  - every pointer is DEVICE memory owned by the caller (torch tensors' data_ptr()); the library
 *    never allocates user-visible memory and keeps no global state besides the last-error string;
 *  - fp32 row-major, int32 indices; `ld*` = leading dimension (elements between consecutive rows);
 *  - every call is asynchronous on `stream` (a cudaStream_t passed as void*), graph-capturable;
 *  - return value: 0 = OK, <0 = error, message via mi_last_error().
 */
#ifndef MATINVENT_B200_H
#define MATINVENT_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define MI_OK 0
#define MI_ERR_ARG (-1)
#define MI_ERR_CUDA (-2)
#define MI_ERR_UNSUPPORTED (-3)

typedef void* mi_stream_t;

const char* mi_last_error(void);
int mi_version(void);
/* fills sm_count / compute capability of the current device */
int mi_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ---------------------------------------------------------------- dense blocks (nn.Linear, ATen addmm)
 * C[M,N] = epilogue(alpha * op(A)[M,K] * op(B)[K,N]).
 *   transA = 0: A stored [M,K] (lda >= K);  transA = 1: A stored [K,M] (lda >= M)
 *   transB = 0: B stored [K,N] (ldb >= N);  transB = 1: B stored [N,K] (ldb >= K)  (nn.Linear weight)
 * Epilogue, per element (m,n), in this order:
 *   v = alpha*acc (+ bias[n]) (+ g1[idx1[m]][n]) (+ g2[idx2[m]][n]) (+ g3[idx3[m]][n]) (+ beta*C_old)
 *   if z_out: z_out[m][n] = v
 *   act == MI_ACT_SILU : v = silu(v);   act == MI_ACT_DSILU : v = v * silu'(z_in[m][n])
 *   if resid: v += resid[m][n];   C[m][n] = v
 * splitk > 1 partitions K over gridDim.z and accumulates with atomics; it requires a plain epilogue
 * (only alpha, and beta == 1 semantics: C += result).
 * Replaces: nn.Linear / torch.cat / gather / SiLU chains of models/diffcsp/cspnet.py:45-54,59-82,264-294
 * and their autograd backward. */
#define MI_ACT_NONE 0
#define MI_ACT_SILU 1
#define MI_ACT_DSILU 2

typedef struct {
    const float* bias;
    const float* g1; const int* g1_idx; int g1_ld;
    const float* g2; const int* g2_idx; int g2_ld;
    const float* g3; const int* g3_idx; int g3_ld;
    float* z_out; int z_ld;
    const float* z_in; int zin_ld;
    const float* resid; int resid_ld;
    int act;
    float alpha;
    float beta;
    int splitk;
    float* amax_out;      /* optional [M]: row-wise max |C[m][:]| (atomic max; caller zeroes it first) */
    const float* a_amax;  /* optional [M], mi_tc_gemm only: row-wise max |A[m][:]| from the producer of A; rows
                             are rescaled by a power of two into fp16 range before the split (exact) and the
                             result rows scaled back, so the tensor-core path keeps fp32 dynamic range */
    const float* col_scale; /* optional [N], mi_tc_gemm only: the accumulator column n is multiplied by col_scale[n]
                             (together with alpha) before bias / gathers: undoes the per-row power-of-two scales
                             mi_f16_split_rows applied to W (mi_sgemm rejects it) */
    /* Fused scatter-mean (mi_tc_gemm, MI_TC_MERGED, plain or z_out epilogue; mi_sgemm rejects it): instead of storing
     * C, rows are reduced by segment:  scat_out[scat_idx[m]][n] += scat_w[m] * v(m, n)   for every row m.  Rows of one
     * segment must be consecutive (CSR order over the destination), scat_w[m] = 1 / (rows in the segment of m) gives
     * the mean of torch_scatter.scatter(reduce='mean') (models/diffcsp/cspnet.py:79); scat_out [S, scat_ld] must be
     * zeroed by the caller; C may be NULL.  scat_amax (nullable, [S], zeroed by the caller) receives max |v| over the
     * segment's rows and all columns: an upper bound of the row maximum of scat_out, usable as the next GEMM's a_amax. */
    float* scat_out; int scat_ld;
    const int* scat_idx;
    const float* scat_w;
    float* scat_amax;
} mi_epilogue_t;

int mi_sgemm(int transA, int transB, int M, int N, int K, const float* A, int lda, const float* B,
             int ldb, float* C, int ldc, const mi_epilogue_t* epi, mi_stream_t stream);

/* Tensor-core variant for the forward dense blocks (A [M,K] and W [N,K] both K-contiguous, i.e.
 * transA = 0, transB = 1): FP32-grade product by split-precision FP16 on tcgen05 (x = x_hi + x_lo, three
 * MMAs per k-slice, TMA-staged swizzled operand tiles, TMEM accumulators), same epilogue contract as
 * mi_sgemm (no beta; MI_ACT_DSILU only without gathers / z_out / g3 / resid; epi->splitk > 1 cuts K into that many
 * parts and ADDS alpha * A W^T to C with 16-byte reductions: plain epilogue only).  W_hi / W_lo are fp16 arrays with the layout of W, produced by mi_f16_split
 * (elementwise, once per weight update).  Requires lda % 4 == 0, ldw % 8 == 0, 16-byte aligned A, W_hi, W_lo.
 *
 * Two operand formats (flags):
 *   0             hi = fp16(x), lo = fp16((x - hi) * 2^11); 128-column tiles with a separate, 2^11-scaled
 *                 correction accumulator (lowest rounding error).  Needs |W| < 65504; A rows larger than
 *                 2^15 need epi->a_amax (see mi_epilogue_t).
 *   MI_TC_MERGED  hi = fp16(s x), lo = fp16(s x - hi) with s a power of two that brings max |x| into
 *                 [2^14, 2^15) (per output row of W: mi_f16_split_rows, undone by epi->col_scale; per row of A
 *                 via epi->a_amax, which is then mandatory): all three products go into ONE accumulator, so a
 *                 128x256 tile double-buffers in TMEM; half the operand traffic per flop, about 2.5x the
 *                 (still ~1e-6 relative) rounding error.
 * mi_f16_split(w, hi, lo, n, scale, lo_scale): hi = fp16(scale * w), lo = fp16((scale * w - hi) * lo_scale);
 * (1, 2048) is the format of flags = 0, (s, 1) the merged one with a caller-chosen s (folded into epi->alpha).
 * mi_f16_split_rows(w [rows, cols] ld, hi, lo (same ld), inv_scale [rows]): merged format with one scale per row,
 * chosen on the device (no host synchronisation): s_r = 2^(14 - exponent(max_c |w[r][c]|)), inv_scale[r] = 1 / s_r. */
#define MI_TC_MERGED 1
int mi_f16_split(const float* w, void* hi, void* lo, long long n, float scale, float lo_scale, mi_stream_t stream);
int mi_f16_split_rows(const float* w, int rows, int cols, int ld, void* hi, void* lo, float* inv_scale,
                      mi_stream_t stream);
int mi_tc_gemm(int M, int N, int K, const float* A, int lda, const void* W_hi, const void* W_lo, int ldw,
               float* C, int ldc, const mi_epilogue_t* epi, int flags, mi_stream_t stream);
/* Same, with A already split by its producer into fp16 (hi, lo) arrays of leading dimension lda
 * (lda % 8 == 0) in the format `flags` names: the operand tiles are loaded by TMA directly, no in-kernel split
 * and no row rescaling (mi_edge_fourier emits this form of the Fourier basis, which is bounded by 1; in the
 * merged format it is scaled by 2^14 and the caller folds 2^-14 into epi->alpha). */
int mi_tc_gemm_presplit(int M, int N, int K, const void* A_hi, const void* A_lo, int lda, const void* W_hi,
                        const void* W_lo, int ldw, float* C, int ldc, const mi_epilogue_t* epi, int flags,
                        mi_stream_t stream);

/* The node-level chain of a CSPNet layer boundary in ONE launch (inference; hidden_dim H = 512): thread-block clusters of
 * four CTAs own 128 rows each, phases separated by cluster barriers, every intermediate operand written pre-split (fp16
 * hi / 2^11-scaled lo, [M, H], row stride H) by the producing epilogue (csrc/mi_node.cu):
 *   prologue xs = split(agg) with the row scales amax_agg calls for; zero_agg != 0: agg is zeroed afterwards (the next fused
 *            scatter-mean's destination)
 *   phase 0  ys = split(silu(agg W_b^T + R + bn1))      node_mlp.0 (cspnet.py:77-82); W_b = node_mlp.0.weight[:, H:] as fp16
 *                                                   (hi, lo) with row stride ld_wb, R = LN(h) node_mlp.0.weight[:, :H]^T
 *                                                   (third block of the previous [P'|Q|R] GEMM, whose row maxima are amax_pqr)
 *   phase 1  h   = h_in + silu(an1 W_2^T + bn2)     node_mlp.2 + residual (cspnet.py:82, 91); h may alias h_in;
 *            xs  = split(LN(h; ln_g, ln_b))          the NEXT layer's LayerNorm (cspnet.py:86-88)
 *   phase 2  pqr = LN(h) W_pqr^T + cb[node_graph]   its per-node GEMM, W_pqr [3H, H]; amax_next [M] (zeroed by the caller)
 *            receives the row maxima of pqr
 * n_phases = 2 stops after phase 1's residual (last layer).  bounds: 3 device floats {max_j ||W_b[j]||_1, max |bn1|,
 * sqrt(H) max |ln_g| + max |ln_b|} (each rounded UP): the row scales of ys and of the LN operand come from these a-priori
 * bounds of the row maxima, not from the maxima.  Replaces four launches (mi_tc_gemm x3 + mi_layernorm_fwd_split).
 * The library runs one of two kernels with identical results up to the FP32-grade rounding of the products: rows on the MMA's
 * N side while the row blocks it forms are at most 64 rows, rows on the M side above (MI_NODE_T=0 / 1 forces one form,
 * MI_NODE_ROWS the rows per block: developer switches, read once per process). */
int mi_node_chain(int M, int H, int n_phases, float* agg, int ld_agg, const float* amax_agg, int zero_agg, void* xs_hi,
                  void* xs_lo, void* ys_hi, void* ys_lo, const void* wb_hi, const void* wb_lo, int ld_wb, const float* bn1,
                  const float* R, int ld_r, const float* amax_pqr, const float* bounds, const void* w2_hi, const void* w2_lo,
                  const float* bn2, const float* h_in, int ld_hin, float* h, int ld_h, const float* ln_g, const float* ln_b,
                  float ln_eps, const void* wpqr_hi, const void* wpqr_lo, const float* cb, int ld_cb, const int* node_graph,
                  float* pqr, int ld_pqr, float* amax_next, mi_stream_t stream);

/* The two per-edge blocks of a CSPNet layer on CTA pairs (tcgen05.mma.cta_group::2, 256 x 256 tiles, both operands
 * pre-split and staged by TMA; csrc/mi_edge.cu), merged operand format (MI_TC_MERGED).  Inference path of
 * models/diffcsp/cspnet.py:59-79; N % 256 == 0.
 *   mi_edge_block1:  a = silu(alpha * (phi W^T) * col_scale[n] + P[src[e]][n] + Q[dst[e]][n]), written as the fp16 pair
 *       a_hi = fp16(s_e a), a_lo = fp16(s_e a - a_hi) [E, ld_a] with s_e the power of two that brings
 *       a_bound[e] = wf_bound[0] + amax_pq[src[e]] + amax_pq[dst[e]] into [2^14, 2^15); a_bound [E] is written too (the
 *       consumer's epi->a_amax).  phi_hi / phi_lo: mi_edge_fourier's merged-format pair (scaled 2^14: alpha = 2^-14);
 *       w_hi / w_lo / col_scale: mi_f16_split_rows of W [N, K]; amax_pq [nodes]: upper bounds of max |P[i][:]|, |Q[i][:]|;
 *       wf_bound: one device float >= max |phi W^T| (sqrt(K / 2) max_j ||W[j]||_2 for the Fourier basis).
 *   mi_edge_block2:  scat_out[scat_idx[e]][n] += scat_w[e] * silu((a W^T)[e][n] * col_scale[n] + bias[n]) with a given as
 *       block 1 wrote it; scat_out (zeroed by the caller), scat_idx, scat_w, scat_amax as in mi_epilogue_t. */
int mi_edge_block1(int E, int N, int K, const void* phi_hi, const void* phi_lo, int ld_phi, const void* w_hi, const void* w_lo,
                   int ld_w, const float* col_scale, float alpha, const float* P, const float* Q, int ld_pq, const int* src,
                   const int* dst, const float* amax_pq, const float* wf_bound, void* a_hi, void* a_lo, int ld_a,
                   float* a_bound, mi_stream_t stream);
int mi_edge_block2(int E, int N, int K, const void* a_hi, const void* a_lo, int ld_a, const float* a_bound, const void* w_hi,
                   const void* w_lo, int ld_w, const float* col_scale, const float* bias, float* scat_out, int scat_ld,
                   const int* scat_idx, const float* scat_w, float* scat_amax, mi_stream_t stream);

/* Transposes that put the weight-gradient GEMM dW[N_out, K_in] += dY^T X (a sum over tens of thousands of edge or node
 * rows) into the K-contiguous form of mi_tc_gemm: dY^T is its fp32 A operand (row maxima = column maxima of dY), X^T
 * its pre-split merged-format W operand with one power-of-two scale per row (= per column of X).
 *   mi_transpose_amax : XT[c][m] = X[m][c] (XT nullable) and col_amax[c] = max(col_amax[c], max_m |X[m][c]|) (nullable)
 *   mi_transpose_split: hi/lo[c][m] = fp16 split of s_c X[m][c], s_c from col_amax[c]; inv_scale[c] = 1 / s_c
 * X [M, C] with row stride ldx; outputs [C, M] with row stride ldt >= M (elements m >= M are not written: zero them once). */
int mi_transpose_amax(const float* X, int ldx, int M, int C, float* XT, int ldt, float* col_amax, mi_stream_t stream);
int mi_transpose_split(const float* X, int ldx, int M, int C, const float* col_amax, void* hi, void* lo, int ldt,
                       float* inv_scale, mi_stream_t stream);

/* ---------------------------------------------------------------- graph construction
 * Fully-connected intra-crystal edges, row-major by (i, j) incl. i == j
 * (models/diffcsp/cspnet.py:238-242: block_diag + dense_to_sparse restated arithmetically).
 *   node_off [B+1], edge_off [B+1] (prefix sums of n_b and n_b^2)
 * out: edge_src, edge_dst, edge_graph [E]; seg_ptr [N+1] (CSR over edge_src: edges of node i are
 *      seg_ptr[i]..seg_ptr[i+1]); dst_ptr [N+1] + dst_perm [E] (CSR over edge_dst: edge ids whose
 *      dst is node j); node_graph [N]. */
int mi_fc_edges(const int* node_off, const int* edge_off, int B, int N, int E, int* edge_src,
                int* edge_dst, int* edge_graph, int* seg_ptr, int* dst_ptr, int* dst_perm,
                int* node_graph, mi_stream_t stream);

/* frac_diff[e] = (x[dst] - x[src]) mod 1 (cspnet.py:242) when cell_off == NULL, else
 * -( -(x[dst] - x[src] + cell_off[e]) ) restated as x[dst]-x[src]+cell_off (knn, cspnet.py:252-257),
 * followed by the Fourier basis Phi[e] = [sin(d_c * 2 pi k)]_{c<3,k<F} || [cos(...)] (cspnet.py:12-24),
 * evaluated with the reference's fp32 arithmetic (arg = d * float(2 pi k)).
 * frac_diff (nullable) [E,3]; phi (nullable) [E, 6F] fp32 with leading dimension ld_phi; phi_hi / phi_lo
 * (nullable pair) the same matrix split for mi_tc_gemm_presplit (same ld): phi_hi = fp16(op_scale * Phi),
 * phi_lo = fp16((op_scale * Phi - phi_hi) * lo_scale) — (1, 2048) or, merged format, (2^14, 1). */
int mi_edge_fourier(const float* x, const int* edge_src, const int* edge_dst, const float* cell_off,
                    int E, int F, float* frac_diff, float* phi, int ld_phi, void* phi_hi, void* phi_lo,
                    float op_scale, float lo_scale, mi_stream_t stream);

/* ---------------------------------------------------------------- segment reductions
 * out[s][:] = scale_s * sum_{k in [ptr[s], ptr[s+1])} X[perm ? perm[k] : k][:]
 *   mean != 0: scale_s = 1 / max(count, 1) (torch_scatter.scatter(reduce='mean'), cspnet.py:79,281)
 *   mean == 0: plain sum.  accumulate != 0: out += result.
 * H must be a multiple of 4 and rows 16-byte aligned.  amax_out (nullable, [S]) receives max |out[s][:]|
 * (for the row rescaling of mi_tc_gemm).  This is the edge-scatter roofline kernel (the inference path forms the
 * scatter-mean of the second per-edge block in that GEMM's epilogue, mi_edge_block2; this kernel serves the knn graphs
 * and every segment sum of the backward). */
int mi_segment_reduce(const float* X, int ldx, const int* ptr, const int* perm, float* out, int ldo,
                      int S, int H, int mean, int accumulate, float* amax_out, mi_stream_t stream);

/* dX[e][:] = dOut[idx[e]][:] * (inv_count ? 1/max(cnt(idx[e]),1) : 1) * (z ? silu'(z[e][:]) : 1)
 * (backward of segment mean + SiLU).  idx nullable -> identity; ptr = CSR used for counts (nullable ->
 * no scaling).  amax_out (nullable, [E], caller zeroes it) receives max |dX[e][:]| for the row rescaling of the
 * tensor-core GEMM that consumes dX. */
int mi_gather_rows_dsilu(const float* dOut, int ldd, const int* idx, const int* ptr, const float* z,
                         int ldz, float* dX, int ldx, int E, int H, float* amax_out, mi_stream_t stream);

/* out[r] = max_c |X[r][c]|: row maxima of an operand whose producer does not report them (the atom-type state of the
 * sampler before the embedding GEMM), for epi->a_amax */
int mi_row_amax(const float* X, int ldx, int rows, int cols, float* out, mi_stream_t stream);
/* out[n] (+)= sum_m X[m][n]  (bias gradients) */
int mi_colsum(const float* X, int ldx, int M, int N, float* out, int accumulate, mi_stream_t stream);

/* ---------------------------------------------------------------- LayerNorm (cspnet.py:57,86-88,142,277)
 * amax_out (nullable, [rows]): atomic max of |y[row][:]| (row rescaling of the tensor-core GEMM reading y) */
int mi_layernorm_fwd(const float* x, int ldx, const float* gamma, const float* beta, float* y, int ldy,
                     float* mean, float* rstd, int rows, int H, float eps, float* amax_out, mi_stream_t stream);
/* LayerNorm feeding a tensor-core GEMM directly: y is written as the pre-split fp16 operand pair of
 * mi_tc_gemm_presplit (flags = 0 format) scaled per row by the power of two that GEMM derives from epi->a_amax, which
 * this call stores into amax [rows] (pass it as epi->a_amax: the GEMM then scales the result rows back).  y (fp32,
 * nullable) is also written when given (training keeps it for the backward); zero_out (nullable): zero_cols floats per
 * row are zeroed (the destination of the fused scatter-mean that follows in the layer).  H <= 1024. */
int mi_layernorm_fwd_split(const float* x, int ldx, const float* gamma, const float* beta, float* y, int ldy,
                           void* y_hi, void* y_lo, int ldh, float* amax, float* zero_out, int ldz, int zero_cols,
                           float* mean, float* rstd, int rows, int H, float eps, mi_stream_t stream);
/* dx (+)= LN backward; dgamma/dbeta += (atomics) */
int mi_layernorm_bwd(const float* dy, int lddy, const float* x, int ldx, const float* gamma,
                     const float* mean, const float* rstd, float* dx, int lddx, int accumulate_dx,
                     float* dgamma, float* dbeta, int rows, int H, mi_stream_t stream);

/* ---------------------------------------------------------------- small per-crystal pieces
 * ips[b] = vec(L_b L_b^T)  (cspnet.py:67-72) */
int mi_lattice_ip(const float* L, float* ips, int B, mi_stream_t stream);
/* out[b][:] = bias + vec(L_b L_b^T) W^T with W [H,9]: the per-crystal term C_b of the split first edge linear
 * (cspnet.py:45,67-72).  n_sets weight sets (the layers of the network: the lattices are the same for all of them)
 * in one launch: set s uses W + s * w_stride, bias + s * bias_stride and writes out + s * out_stride. */
int mi_lattice_linear(const float* L, const float* W, const float* bias, float* out, int ldo, int B, int H,
                      int n_sets, long long w_stride, long long bias_stride, long long out_stride,
                      mi_stream_t stream);
/* The output heads of CSPNet at inference in one launch, one CTA per crystal (cspnet.py:276-294):
 *   hf = LayerNorm(h; ln_g, ln_b) (ln_g == NULL: hf = h);  pred_x [N,3] = hf coord_w^T;  pred_a [N,A] = hf type_w^T + type_b;
 *   pred_l [B,3,3] = lattice_w (mean over the crystal's atoms of hf), multiplied by L[b] from the right when ip != 0.
 * Any of pred_x / pred_a / pred_l may be NULL (the corrector evaluation only needs pred_x).  FP32 CUDA-core arithmetic;
 * replaces mi_layernorm_fwd + 3 GEMM launches + mi_segment_reduce + mi_bmm3.  H in {128, 256, 512, 1024}. */
int mi_output_heads(const float* h, int ldh, const int* node_off, int B, int H, const float* ln_g, const float* ln_b, float eps,
                    const float* coord_w, float* pred_x, const float* type_w, const float* type_b, int A, float* pred_a,
                    const float* lattice_w, const float* L, int ip, float* pred_l, mi_stream_t stream);
/* out[b] = A[b] (3x3) @ L[b] (3x3)   (cspnet.py:288-289); transL != 0 -> A[b] @ L[b]^T (its backward) */
int mi_bmm3(const float* A, const float* L, float* out, int B, int transL, mi_stream_t stream);
/* SinusoidalTimeEmbeddings (diffusion.py:53-66): out[b] = [sin(t_b f_k) || cos(t_b f_k)], dim even;
 * freq [dim/2] = exp(-k ln(1e4)/(dim/2-1)) is a host-built table (bit-identical to the reference's). */
int mi_time_embed(const int* t, const float* freq, int B, int dim, float* out, mi_stream_t stream);
/* lattice_params_to_matrix_torch (utils.py:68-96); lengths/angles [B,3] (degrees) -> L [B,3,3] */
int mi_lattice_params_to_matrix(const float* lengths, const float* angles, float* L, int B,
                                mi_stream_t stream);
/* lattices_to_params_shape (sample.py:103-114) + argmax(atom_types)+1 (sample.py:182) */
int mi_lattice_matrix_to_params(const float* L, float* lengths, float* angles, int B,
                                mi_stream_t stream);
int mi_argmax_rows(const float* a, int lda, int rows, int cols, int add, int* out, mi_stream_t stream);

/* ---------------------------------------------------------------- reverse diffusion updates
 * Scalars of step t (diffusion.py:300-307,324-325,341-343) come from a device table built once on the
 * host in fp32: coef[t] = {sqrt(sigma_norm_t), step_c, std_c, step_p, std_p, c0, c1, sigma^beta_t}.
 * The row is selected by *t_dev when t_dev != NULL (so ONE captured CUDA graph serves every step),
 * else by t_host.  Noise pointers may be NULL (treated as zeros: t == 1).
 * corrector (diffusion.py:327-330): x_half = x - step_c * (pred_x * sqrt_sn) + std_c * z_x */
int mi_reverse_corrector(const float* x, const float* pred_x, const float* z_x, float* x_half, int N,
                         const float* coef, const int* t_dev, int t_host, mi_stream_t stream);
/* predictor (diffusion.py:345-351,386): x = ((x_half - step_p*pred_x*sqrt_sn + std_p*z_x) % 1) % 1;
 * l = c0*(l - c1*pred_l) + sig*z_l ; a = c0*(a - c1*pred_a) + sig*z_a  (in place on l, a) */
int mi_reverse_predictor(const float* x_half, const float* pred_x, const float* z_x, float* x, int N,
                         float* l, const float* pred_l, const float* z_l, int B, float* a,
                         const float* pred_a, const float* z_a, int A, const float* coef,
                         const int* t_dev, int t_host, mi_stream_t stream);
/* temb[b][:] = ttab[*t_dev][:] for every crystal (diffusion.py:297-298); *t_dev -= 1 */
int mi_sampler_step_begin(const int* t_dev, const float* ttab, float* temb, int B, int T,
                          mi_stream_t stream);
int mi_sampler_step_end(int* t_dev, mi_stream_t stream);

/* ---------------------------------------------------------------- forward noising + losses
 * add_noise (diffusion.py:81-119) for one integer time shared by the batch:
 *   l_t = c0*L0 + c1*z_l ; x_t = (x0 + sigma*z_x) % 1 ; a_t = c0*onehot(Z-1) + c1*z_a ;
 *   tar_x = d_log_p_wrapped_normal(sigma*z_x, sigma) / sqrt(sigma_norm)  (scheduler.py:39-43)
 * coef[t] = {sqrt(abar_t), sqrt(1-abar_t), sigma_t, sqrt(sigma_norm_t)} (device table, row = *t_dev or t_host) */
int mi_add_noise(const float* L0, const float* x0, const int* Z, const float* z_l, const float* z_x,
                 const float* z_a, int B, int N, int A, const float* coef, const int* t_dev, int t_host,
                 float* l_t, float* x_t, float* a_t, float* tar_x, mi_stream_t stream);

/* Per-crystal denoising loss (diffusion.py:121-138), KL proxy (diffusion.py:140-149) and the
 * reward-weighted objective of MatInvent.ft_step (pipeline/mat_invent.py:152-163), with the
 * gradient of   J = scale * sum_b [ w_loss[b]*loss_b + w_kl[b]*kl_b ]   w.r.t. the agent's
 * predictions.  prior_* may be NULL (kl = 0).  d_* may be NULL (forward only).
 * Host passes w_loss = reward, w_kl = sigma*(1.1-reward), scale = 1/(B_global*accum_steps), or the
 * upstream autograd gradient in the plugin path.  stats (nullable, 2 floats) accumulates
 * sum_b w_loss*loss_b and sum_b w_kl*kl_b for the ft_step logs (pipeline/mat_invent.py:168-170). */
int mi_rl_loss(const float* pred_l, const float* pred_x, const float* pred_a, const float* tgt_l,
               const float* tgt_x, const float* tgt_a, const float* prior_l, const float* prior_x,
               const float* prior_a, const int* node_off, int B, int A, float cost_l, float cost_x,
               float cost_a, const float* w_loss, const float* w_kl, float scale, float* loss,
               float* kl, float* d_l, float* d_x, float* d_a, float* stats, mi_stream_t stream);

/* ---------------------------------------------------------------- optimiser
 * torch.optim.Adam defaults (pipeline/mat_invent.py:136,166): flat fp32 buffers,
 *   m = m + (g-m)*(1-b1) ; v = b2*v + (1-b2)*g*g ;
 *   p -= (lr/(1-b1^t)) * m / (sqrt(v)/sqrt(1-b2^t) + eps);  g is multiplied by grad_scale first and
 * zeroed afterwards when zero_grad != 0. */
int mi_adam_step(float* p, float* g, float* m, float* v, long long n, double lr, double b1, double b2,
                 double eps, int step, float grad_scale, int zero_grad, mi_stream_t stream);

/* ---------------------------------------------------------------- RNG
 * Philox4x32-10 + Box-Muller standard normals / uniforms in [0,1): element i of the call uses counter
 * (offset + i/4), key = seed.  For CUDA-graph replay, `offset_dev` (nullable) is a device u64 added to
 * `offset` and, when advance != 0, incremented by ceil(n/4) by the kernel afterwards. */
int mi_philox_normal(float* out, long long n, unsigned long long seed, unsigned long long offset,
                     unsigned long long* offset_dev, int advance, mi_stream_t stream);
int mi_philox_uniform(float* out, long long n, unsigned long long seed, unsigned long long offset,
                      unsigned long long* offset_dev, int advance, mi_stream_t stream);

/* ---------------------------------------------------------------- periodic neighbour list
 * radius_graph_pbc + get_max_neighbors_mask (models/diffcsp/utils.py:335-514, 517-601) followed by
 * reorder_symmetric_edges (cspnet.py:159-234) and the sign flip of gen_edges (cspnet.py:252-257),
 * emitted sorted by source node (the network is invariant to edge order).
 * One CTA per crystal (n_b <= max_n <= 128).  Capacity: each node holds at most `cap` edges; the true
 * count is written to deg[N] (and to overflow[0] != 0 if any node exceeded cap).
 * out: edge_dst [N*cap], cell_off [N*cap,3] (as float), deg [N]; compact with mi_compact_edges. */
int mi_radius_graph_pbc(const float* x, const float* L, const int* node_off, int B, int N, int max_n,
                        int max_neighbors, int cap, int* edge_dst, float* cell_off, int* deg,
                        int* overflow, mi_stream_t stream);
/* seg_ptr = exclusive scan of deg (single CTA), then gather padded rows into CSR arrays */
int mi_compact_edges(const int* deg, int N, int cap, const int* edge_dst_pad, const float* cell_pad,
                     const int* node_graph, int* seg_ptr, int* edge_src, int* edge_dst,
                     int* edge_graph, float* cell_off, int E_cap, mi_stream_t stream);
/* CSR over destinations from (edge_dst) : dst_ptr [N+1], dst_perm [E] (E read from seg_ptr[N]) */
int mi_build_dst_csr(const int* seg_ptr, const int* edge_dst, int N, int E_cap, int* dst_ptr,
                     int* dst_perm, int* work, mi_stream_t stream);

/* ---------------------------------------------------------------- device replay buffer
 * memory/replay_buffer.py:32-73 on padded SoA rows: given keys (64-bit reduced-composition hash) and
 * rewards of `n` candidate rows (old rows first), compute the kept row order:
 * sort by reward desc (stable), drop later duplicates of a key, keep the first `buffer_size`, keep
 * reward > cutoff.  Rewards are float64 like the reference's pandas column.  out_idx [n] (first *out_count entries
 * valid).  Single CTA: n <= 16384 candidate rows (buffer + one iteration's top-k; the caller pre-selects above that).
 * Rows are identified by the 64-bit key alone (no collision check: 2^-64 per pair). */
int mi_replay_select(const unsigned long long* keys, const double* rewards, int n, int buffer_size,
                     double cutoff, int* out_idx, int* out_count, mi_stream_t stream);
/* key[b] = hash of the gcd-reduced element-count vector of crystal b (Z in 1..100)
 * (pymatgen reduced_formula equivalence class, memory/replay_buffer.py:38) */
int mi_composition_key(const int* Z, const int* node_off, int B, unsigned long long* keys,
                       mi_stream_t stream);

/* ---------------------------------------------------------------- post-sampling pipeline
 * Validity pre-filter of sampled crystals, one warp per crystal; replaces the per-structure Python / mp.Pool loop of
 * pipeline/filters/opt_filter.py:38-63 (`invalid_filter`).  frac [N,3], L [B,3,3] (rows = lattice vectors),
 * lengths [B,3], node_off [B+1].  mask[b] bit 0 = max(a,b,c) < max_len (the in-tree rule, opt_filter.py:53-55);
 * bit 1 = min periodic interatomic distance (27 images) >= min_dist, |det L| >= min_vol, max(a,b,c) <= hard_len
 * (mattergen `structure_validity`, opt_filter.py:51: un-vendored, restated from its published definition —
 * parity unpinned).  dmin (nullable, [B]) receives the minimum distance. */
int mi_validity_prefilter(const float* frac, const float* L, const float* lengths, const int* node_off, int B,
                          float max_len, float min_dist, float min_vol, float hard_len, int* mask, float* dmin,
                          mi_stream_t stream);
/* Composition-level rewards, one warp per crystal: prop[p][b] = sum_el w_el * tables[p][el] with w = mass fraction
 * (modes[p] = 0) or atomic fraction (1) — the form of rewards/calculators/pymatgen/calc.py:24-45,57-92 (hhi, price,
 * crustal abundance); NaN table entry of a present element => failed sample (calc.py:63-70).  Then
 * rewards/reward.py:51-115 in float64: nan_to_num, linear_scaling (targets[p] = 0 ascending, 1 descending, 2 target
 * value tval[p]; minv/maxv), reduce (0 mean, 1 min, 2 weighted sum), failed -> reward 0.
 * Z [N] (1..100), node_off [B+1], tables [P,128] and mass [128] device doubles indexed by atomic number; the
 * per-property configuration arrays are HOST arrays of length P (<= 8).  Out: props [P,B], rewards [B], failed [B]. */
int mi_composition_reward(const int* Z, const int* node_off, int B, const double* tables, const double* mass, int P,
                          const int* modes, const int* targets, const double* minv, const double* maxv,
                          const double* tval, const double* weight, int reduce, double* props, double* rewards,
                          int* failed, mi_stream_t stream);

/* MatterGen adapter (models/mattergen/loss.py:63-73): out[b] = sum_k weights[k] * f_k[b] over F <= 4 per-field
 * per-sample loss vectors, accumulated in the given order (weights: HOST array of F floats). */
int mi_weighted_field_sum(int B, int F, const float* f0, const float* f1, const float* f2, const float* f3,
                          const float* weights_host, float* out, mi_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif
