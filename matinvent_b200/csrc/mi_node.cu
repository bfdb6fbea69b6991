// The node-level chain of a CSPNet layer boundary in ONE launch (inference):
//
//   phase 0   an1   = silu(agg W_b^T + R + b_n1)             node_mlp.0 on [LN(h) | agg]; R = LN(h) W_a^T comes from the
//                                                            previous P|Q|R GEMM (cspnet.py:77-82, forward_graph "node path")
//   phase 1   h    += silu(an1 W_n2^T + b_n2)                node_mlp.2 + residual (cspnet.py:82, 91)
//   phase 2   [P'|Q|R] = LN(h) [W_hi; W_hj; W_a]^T + [C_b|0|0]   the NEXT layer's LayerNorm (cspnet.py:86-88) applied on the
//                                                            fly to the A operand, and its per-node GEMM
//
// As separate launches these are four latency-bound kernels of one 128-row tile per CTA (~15-23 us each for 3-8 us of
// tensor work: pipeline fill, an exposed epilogue, launch ramp and tail every time).  The GEMMs are row-wise, so the only
// dependency between CTAs is across the column tiles of the same 128 rows: a thread-block CLUSTER of four CTAs owns one row
// block (each CTA 128 of the 512 output columns, 384 of the 1536 in phase 2), the phases are separated by cluster barriers
// and the intermediate activations go through global memory (L2) to the peers' TMA loads.  One launch, one pipeline fill.
//
// Same arithmetic as mi_tc.cu's two-accumulator format (split-precision FP16 x3 on tcgen05, rows rescaled by powers of
// two from their maxima): w0 TMA producer, w1 MMA issuer, w2-7 operand split (phase 2: LayerNorm statistics + affine on
// the fly), w8-15 epilogue.  H = 512 only (four 128-column tiles); other sizes use the separate kernels.
#include <cooperative_groups.h>
#include <stdlib.h>

#include "mi_tc_common.cuh"

using namespace mi_tc;

namespace {

constexpr int TM = 128, TN = 128, TK = 32;
constexpr int R = 3, S = 4;                                  // raw fp32 A ring, operand ring
constexpr int A_RAW = TM * TK * 4, A_H = TM * TK * 2, W_H = TN * TK * 2, OPB = 2 * A_H + 2 * W_H;
constexpr int EPI_WARPS = 8, EPI_WARP0 = 8, THREADS = 512, SPLIT_THREADS = 192, SPLIT_WARPS = 6;
constexpr int EP = 34;
constexpr int RAW_BYTES = R * A_RAW, OP_BYTES = S * OPB, EBUF_BYTES = EPI_WARPS * 32 * EP * 4, BAR_BYTES = 256;
constexpr int HMAX = 512;
constexpr int OFF_EBUF = RAW_BYTES + OP_BYTES, OFF_BAR = OFF_EBUF + EBUF_BYTES, OFF_REXP = OFF_BAR + BAR_BYTES;
constexpr int OFF_LNMEAN = OFF_REXP + 128, OFF_LNRSTD = OFF_LNMEAN + 512, OFF_LNEXP = OFF_LNRSTD + 512;
constexpr int OFF_GAMMA = OFF_LNEXP + 128, OFF_BETA = OFF_GAMMA + 4 * HMAX, SMEM_BYTES = OFF_BETA + 4 * HMAX;
static_assert(SMEM_BYTES <= 232448, "does not fit the SM");
static_assert((3 * S + 2 * R + 4) * 8 + 8 <= BAR_BYTES, "barrier block too small");
constexpr uint32_t ACC_COLS = 2 * TN, TMEM_COLS = 2 * ACC_COLS;
constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
constexpr int CLUSTER = 4;

struct Phase {
    int n_tiles;                  // 128-column tiles per CTA (1 or 3)
    int N;                        // output columns in total
    float* C; int ldc;
    const float* bias;
    const float* g1; const int* g1_idx; int g1_ld;       // row gather added before the activation (idx NULL: row m)
    const float* resid; int resid_ld;                    // added after the activation
    int act;
    float* amax_out;              // row maxima of C (atomic max), nullable
    const float* a_amax;          // row maxima of A; NULL in the LayerNorm phase
};
struct Params {
    int M, K, n_phases;
    Phase ph[3];
    // LayerNorm of phase 2: input = phase 1's C
    const float* ln_x; int ln_ldx;
    const float* ln_g; const float* ln_b; float ln_eps;
    float* zero_out; int zero_ld;          // agg rows of this row block are zeroed once phase 0 has consumed them (nullable)
};

__device__ __forceinline__ void cluster_sync_all() {
    // every thread of the four CTAs: global writes of this phase (generic proxy) -> visible to the peers' loads, including
    // their TMA loads (async proxy)
    __threadfence();
    asm volatile("fence.proxy.async;" ::: "memory");
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    asm volatile("fence.proxy.async;" ::: "memory");
}

__global__ void __cluster_dims__(CLUSTER, 1, 1) __launch_bounds__(THREADS, 1)
node_chain_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
                  const __grid_constant__ CUtensorMap mapA2, const __grid_constant__ CUtensorMap mapW0h,
                  const __grid_constant__ CUtensorMap mapW0l, const __grid_constant__ CUtensorMap mapW1h,
                  const __grid_constant__ CUtensorMap mapW1l, const __grid_constant__ CUtensorMap mapW2h,
                  const __grid_constant__ CUtensorMap mapW2l, const Params p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* raw_ring = smem;
    uint8_t* op_ring = smem + RAW_BYTES;
    float* ebuf_all = reinterpret_cast<float*>(smem + OFF_EBUF);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    uint64_t* w_full = bars;
    uint64_t* a_ready = bars + S;
    uint64_t* op_empty = bars + 2 * S;
    uint64_t* raw_full = bars + 3 * S;
    uint64_t* raw_empty = bars + 3 * S + R;
    uint64_t* acc_full = bars + 3 * S + 2 * R;
    uint64_t* acc_empty = acc_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 4);
    int8_t* rexp = reinterpret_cast<int8_t*>(smem + OFF_REXP);
    float* ln_mean = reinterpret_cast<float*>(smem + OFF_LNMEAN);
    float* ln_rstd = reinterpret_cast<float*>(smem + OFF_LNRSTD);
    int8_t* ln_exp = reinterpret_cast<int8_t*>(smem + OFF_LNEXP);
    float* sgamma = reinterpret_cast<float*>(smem + OFF_GAMMA);
    float* sbeta = reinterpret_cast<float*>(smem + OFF_BETA);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int crank = (int)(blockIdx.x % CLUSTER);               // rank in the cluster = column slice
    const int m0 = (int)(blockIdx.x / CLUSTER) * TM;             // the cluster's row block
    const int nkb = (p.K + TK - 1) / TK;

    if (threadIdx.x == 0) {
        if (smem_u32(smem) & 1023u) __trap();
        for (int s = 0; s < S; ++s) {
            mbar_init(&w_full[s], 1);
            mbar_init(&a_ready[s], SPLIT_WARPS);
            mbar_init(&op_empty[s], 1);
        }
        for (int r = 0; r < R; ++r) {
            mbar_init(&raw_full[r], 1);
            mbar_init(&raw_empty[r], SPLIT_WARPS);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&acc_full[b], 1);
            mbar_init(&acc_empty[b], EPI_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA0) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapW0h) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapW0l) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    // pipeline counters run across phases (the mbarrier phases just keep alternating)
    uint32_t g_op = 0;       // k-blocks through the operand ring so far (every role counts its own copy)
    uint32_t g_raw = 0;      // raw tiles issued (producer only)
    uint32_t g_tile = 0;     // tiles so far

    for (int phx = 0; phx < p.n_phases; ++phx) {
        const Phase& ph = p.ph[phx];
        const CUtensorMap* mA = phx == 0 ? &mapA0 : (phx == 1 ? &mapA1 : &mapA2);
        const CUtensorMap* mWh = phx == 0 ? &mapW0h : (phx == 1 ? &mapW1h : &mapW2h);
        const CUtensorMap* mWl = phx == 0 ? &mapW0l : (phx == 1 ? &mapW1l : &mapW2l);
        const bool ln_phase = ph.a_amax == nullptr;
        const int ntl = ph.n_tiles;
        const uint32_t total = (uint32_t)ntl * (uint32_t)nkb;

        if (warp == 0) {
            // ===================== TMA producer =====================
            if (lane == 0) {
                auto issue_raw = [&](uint32_t loc) {
                    const int kb = (int)(loc % (uint32_t)nkb);
                    const uint32_t gi = g_raw++;
                    const int r = (int)(gi % (uint32_t)R);
                    mbar_wait(&raw_empty[r], ((gi / (uint32_t)R) & 1) ^ 1);
                    mbar_expect_tx(&raw_full[r], A_RAW);
                    tma_load_2d(raw_ring + r * A_RAW, mA, &raw_full[r], kb * TK, m0);
                };
                uint32_t a_loc = 0;
                for (; (int)a_loc < R - 1 && a_loc < total; ++a_loc) issue_raw(a_loc);
                for (uint32_t loc = 0; loc < total; ++loc, ++g_op) {
                    const int tl = (int)(loc / (uint32_t)nkb), kb = (int)(loc % (uint32_t)nkb);
                    const int n0 = (crank * ntl + tl) * TN;
                    const int s = (int)(g_op % (uint32_t)S);
                    mbar_wait(&op_empty[s], ((g_op / (uint32_t)S) & 1) ^ 1);
                    uint8_t* st = op_ring + s * OPB;
                    mbar_expect_tx(&w_full[s], 2 * W_H);
                    tma_load_2d(st + 2 * A_H, mWh, &w_full[s], kb * TK, n0);
                    tma_load_2d(st + 2 * A_H + W_H, mWl, &w_full[s], kb * TK, n0);
                    if (a_loc < total) issue_raw(a_loc++);
                }
            } else {
                g_op += total;
            }
        } else if (warp == 1) {
            // ===================== MMA issuer =====================
            if (lane == 0) {
                for (int tl = 0; tl < ntl; ++tl) {
                    const uint32_t gt = g_tile + (uint32_t)tl;
                    const uint32_t ab = gt & 1;
                    const uint32_t acc = tmem_base + ab * ACC_COLS;
                    mbar_wait(&acc_empty[ab], ((gt >> 1) & 1) ^ 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    for (int kb = 0; kb < nkb; ++kb, ++g_op) {
                        const int s = (int)(g_op % (uint32_t)S);
                        const uint32_t par = (g_op / (uint32_t)S) & 1;
                        mbar_wait(&w_full[s], par);
                        mbar_wait(&a_ready[s], par);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const uint32_t st = smem_u32(op_ring + s * OPB);
                        const uint64_t d_ahi = umma_desc(st), d_alo = umma_desc(st + A_H);
                        const uint64_t d_whi = umma_desc(st + 2 * A_H), d_wlo = umma_desc(st + 2 * A_H + W_H);
#pragma unroll
                        for (int k = 0; k < TK / 16; ++k) {
                            const uint64_t adv = (uint64_t)((k * 32) >> 4);
                            umma_f16(acc, d_ahi + adv, d_whi + adv, IDESC, (kb | k) != 0);
                            umma_f16(acc + TN, d_alo + adv, d_whi + adv, IDESC, (uint32_t)((kb | k) != 0));
                            umma_f16(acc + TN, d_ahi + adv, d_wlo + adv, IDESC, 1u);
                        }
                        umma_commit(&op_empty[s]);
                    }
                    umma_commit(&acc_full[ab]);
                }
            } else {
                g_op += total;
            }
        } else if (warp < EPI_WARP0) {
            // ===================== operand split warps (w2..7) =====================
            const int t = threadIdx.x - 64;       // 0..191
            const int sw = warp - 2;
            if (ln_phase) {
                // LayerNorm of the row block: affine parameters to shared memory, then one warp per row: mean, rstd and the
                // power-of-two exponent of max |y| (two passes over registers, the arithmetic of layernorm_fwd_kernel)
                for (int c = t; c < p.K; c += SPLIT_THREADS) {
                    sgamma[c] = __ldg(p.ln_g + c);
                    sbeta[c] = __ldg(p.ln_b + c);
                }
                asm volatile("bar.sync 1, %0;" ::"n"(SPLIT_THREADS) : "memory");
                for (int row = sw; row < TM; row += SPLIT_WARPS) {
                    const int m = m0 + row;
                    float mean = 0.f, rstd = 0.f;
                    int e8 = 0;
                    if (m < p.M) {
                        const float* xr = p.ln_x + (long long)m * p.ln_ldx;
                        float xv[HMAX / 32];
                        float s = 0.f;
#pragma unroll
                        for (int i = 0; i < HMAX / 32; ++i) {
                            const int c = lane + 32 * i;
                            xv[i] = c < p.K ? __ldcg(xr + c) : 0.f;          // written by the peers in phase 1: L2, not L1
                            s += xv[i];
                        }
                        s = mi_warp_sum(s);
                        mean = s / (float)p.K;
                        float v = 0.f;
#pragma unroll
                        for (int i = 0; i < HMAX / 32; ++i) {
                            const float d = (lane + 32 * i < p.K) ? xv[i] - mean : 0.f;
                            v += d * d;
                        }
                        v = mi_warp_sum(v);
                        rstd = 1.0f / sqrtf(v / (float)p.K + p.ln_eps);
                        float mx = 0.f;
#pragma unroll
                        for (int i = 0; i < HMAX / 32; ++i) {
                            const int c = lane + 32 * i;
                            if (c < p.K) mx = fmaxf(mx, fabsf((xv[i] - mean) * rstd * sgamma[c] + sbeta[c]));
                        }
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
                        const int ex = (int)((__float_as_uint(mx) >> 23) & 0xff) - 127;
                        e8 = max(-100, min(ex - 14, 100));
                        // the destination of the next layer's fused scatter-mean: this CTA's column slice of the row
                        if (p.zero_out) {
                            float* zr = p.zero_out + (long long)m * p.zero_ld + crank * (p.K / CLUSTER);
                            for (int c = lane; c < p.K / CLUSTER; c += 32) zr[c] = 0.f;
                        }
                    }
                    if (lane == 0) {
                        ln_mean[row] = mean;
                        ln_rstd[row] = rstd;
                        ln_exp[row] = (int8_t)e8;
                    }
                }
                asm volatile("bar.sync 2, %0;" ::"n"(SPLIT_THREADS + EPI_WARPS * 32) : "memory");      // statistics visible to the epilogue warps too
            }
            for (int tl = 0; tl < ntl; ++tl) {
                if (!ln_phase) {
                    int e8 = 0;
                    const int m = m0 + t;
                    if (t < TM && m < p.M) {
                        const int ex = (int)((__float_as_uint(__ldcg(ph.a_amax + m)) >> 23) & 0xff) - 127;
                        e8 = max(-100, min(ex - 14, 100));
                    }
                    asm volatile("bar.sync 1, %0;" ::"n"(SPLIT_THREADS) : "memory");
                    if (t < TM) rexp[t] = (int8_t)e8;
                    asm volatile("bar.sync 1, %0;" ::"n"(SPLIT_THREADS) : "memory");
                }
                const int8_t* rx = ln_phase ? ln_exp : rexp;
                for (int kb = 0; kb < nkb; ++kb, ++g_op) {
                    const int r = (int)(g_raw % (uint32_t)R), s = (int)(g_op % (uint32_t)S);
                    mbar_wait(&raw_full[r], (g_raw / (uint32_t)R) & 1);
                    const float4* raw = reinterpret_cast<const float4*>(raw_ring + r * A_RAW);
                    uint8_t* hi = op_ring + s * OPB;
                    uint8_t* lo = hi + A_H;
                    constexpr int NV = (TM * TK / 4 + SPLIT_THREADS - 1) / SPLIT_THREADS;
                    float4 v[NV];
                    float sc[NV];
#pragma unroll
                    for (int i = 0; i < NV; ++i) {
                        const int pidx = i * SPLIT_THREADS + t;
                        if (pidx < TM * TK / 4) {
                            v[i] = raw[pidx];
                            const int row = pidx >> 3;
                            sc[i] = __uint_as_float((uint32_t)(127 - (int)rx[row]) << 23);
                            if (ln_phase) {
                                const int k0 = ((pidx & 7) ^ (row & 7)) << 2;
                                const float4 g4 = *reinterpret_cast<const float4*>(sgamma + kb * TK + k0);
                                const float4 b4 = *reinterpret_cast<const float4*>(sbeta + kb * TK + k0);
                                const float mu = ln_mean[row], rs = ln_rstd[row];
                                v[i].x = (v[i].x - mu) * rs * g4.x + b4.x;
                                v[i].y = (v[i].y - mu) * rs * g4.y + b4.y;
                                v[i].z = (v[i].z - mu) * rs * g4.z + b4.z;
                                v[i].w = (v[i].w - mu) * rs * g4.w + b4.w;
                            }
                        }
                    }
                    mbar_wait(&op_empty[s], ((g_op / (uint32_t)S) & 1) ^ 1);
#pragma unroll
                    for (int i = 0; i < NV; ++i) {
                        const int pidx = i * SPLIT_THREADS + t;
                        if (pidx >= TM * TK / 4) break;
                        const int row = pidx >> 3;
                        const int k0 = ((pidx & 7) ^ (row & 7)) << 2;
                        uint2 h, l;
                        split2<0>(v[i].x * sc[i], v[i].y * sc[i], h.x, l.x);
                        split2<0>(v[i].z * sc[i], v[i].w * sc[i], h.y, l.y);
                        const int off = row * 64 + ((((k0 >> 3) ^ (row >> 1)) & 3) << 4) + ((k0 & 7) << 1);
                        *reinterpret_cast<uint2*>(hi + off) = h;
                        *reinterpret_cast<uint2*>(lo + off) = l;
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) {
                        mbar_arrive(&raw_empty[r]);
                        mbar_arrive(&a_ready[s]);
                    }
                    ++g_raw;
                }
            }
        } else {
            // ===================== epilogue warps (w8..15) =====================
            const int q = warp & 3;
            const int cg = (warp - EPI_WARP0) >> 2;               // two column groups of two 32-column chunks
            float* ebuf = ebuf_all + (warp - EPI_WARP0) * (32 * EP);
            if (ln_phase) asm volatile("bar.sync 2, %0;" ::"n"(SPLIT_THREADS + EPI_WARPS * 32) : "memory");
            for (int tl = 0; tl < ntl; ++tl) {
                const uint32_t gt = g_tile + (uint32_t)tl;
                const uint32_t ab = gt & 1;
                const uint32_t acc = tmem_base + ab * ACC_COLS;
                const int n0 = (crank * ntl + tl) * TN;
                const int mrow = m0 + q * 32 + lane;
                int e8 = 0;
                if (mrow < p.M) {
                    if (ln_phase) e8 = (int)ln_exp[q * 32 + lane];
                    else {
                        const int ex = (int)((__float_as_uint(__ldcg(ph.a_amax + mrow)) >> 23) & 0xff) - 127;
                        e8 = max(-100, min(ex - 14, 100));
                    }
                }
                const float rowsc = __uint_as_float((uint32_t)(127 + e8) << 23);
                int i1 = 0;
                if (ph.g1) i1 = mrow < p.M ? (ph.g1_idx ? __ldg(ph.g1_idx + mrow) : mrow) : 0;
                float rowmax[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) rowmax[u] = 0.f;
                mbar_wait(&acc_full[ab], (gt >> 1) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                uint32_t v[32], w[32];
                const uint32_t tbase = acc + ((uint32_t)(q * 32) << 16) + (uint32_t)(cg * 64);
#pragma unroll 1
                for (int cc = 0; cc < 2; ++cc) {
                    const int nb = n0 + (cg * 2 + cc) * 32;
                    const int col4 = (lane & 7) * 4;
                    const int n = nb + col4;
                    tmem_ld32(tbase + (uint32_t)(cc * 32), v);
                    tmem_ld32(tbase + (uint32_t)(cc * 32) + TN, w);
                    // gathered rows of the first half chunk in flight behind the TMEM read
                    float4 ga0[4], ga1[4];
                    auto issue_gathers = [&](int hb, float4 (&ga)[4]) {
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int rr = (hb * 4 + u) * 4 + (lane >> 3);
                            const int r1 = __shfl_sync(0xffffffffu, i1, rr);
                            ga[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (ph.g1) ga[u] = __ldcg(reinterpret_cast<const float4*>(ph.g1 + (long long)r1 * ph.g1_ld + n));
                        }
                    };
                    issue_gathers(0, ga0);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    if (cc == 1) {
                        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&acc_empty[ab]);
                    }
#pragma unroll
                    for (int j = 0; j < 32; j += 2) {
                        const float y0 = fmaf(__uint_as_float(w[j]), LO_UNSCALE, __uint_as_float(v[j]));
                        const float y1 = fmaf(__uint_as_float(w[j + 1]), LO_UNSCALE, __uint_as_float(v[j + 1]));
                        *reinterpret_cast<float2*>(ebuf + lane * EP + j) = make_float2(rowsc * y0, rowsc * y1);
                    }
                    __syncwarp();
                    issue_gathers(1, ga1);
                    float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (ph.bias) bias4 = __ldg(reinterpret_cast<const float4*>(ph.bias + n));
#pragma unroll
                    for (int hb = 0; hb < 2; ++hb) {
                        float4 gr[4];
                        float2 xa[4], xb[4];
                        int mm[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int rr = (hb * 4 + u) * 4 + (lane >> 3);
                            const int m = m0 + q * 32 + rr;
                            mm[u] = m;
                            gr[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (ph.resid && m < p.M) gr[u] = __ldcg(reinterpret_cast<const float4*>(ph.resid + (long long)m * ph.resid_ld + n));
                            xa[u] = *reinterpret_cast<const float2*>(ebuf + rr * EP + col4);
                            xb[u] = *reinterpret_cast<const float2*>(ebuf + rr * EP + col4 + 2);
                        }
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int m = mm[u];
                            if (m >= p.M) continue;
                            const float4 a4 = hb ? ga1[u] : ga0[u];
                            float x[4] = {xa[u].x + bias4.x + a4.x, xa[u].y + bias4.y + a4.y, xb[u].x + bias4.z + a4.z,
                                          xb[u].y + bias4.w + a4.w};
                            if (ph.act == MI_ACT_SILU) {
#pragma unroll
                                for (int v4 = 0; v4 < 4; ++v4) x[v4] = silu_fast(x[v4]);
                            }
                            x[0] += gr[u].x; x[1] += gr[u].y; x[2] += gr[u].z; x[3] += gr[u].w;
                            *reinterpret_cast<float4*>(ph.C + (long long)m * ph.ldc + n) = make_float4(x[0], x[1], x[2], x[3]);
                            rowmax[hb * 4 + u] = fmaxf(rowmax[hb * 4 + u], fmaxf(fmaxf(fabsf(x[0]), fabsf(x[1])), fmaxf(fabsf(x[2]), fabsf(x[3]))));
                        }
                    }
                    __syncwarp();
                }
                if (ph.amax_out) {
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        float rmax = rowmax[u];
                        rmax = fmaxf(rmax, __shfl_xor_sync(0xffffffffu, rmax, 1));
                        rmax = fmaxf(rmax, __shfl_xor_sync(0xffffffffu, rmax, 2));
                        rmax = fmaxf(rmax, __shfl_xor_sync(0xffffffffu, rmax, 4));
                        const int m = m0 + q * 32 + u * 4 + (lane >> 3);
                        if ((lane & 7) == 0 && m < p.M) atomicMax(reinterpret_cast<unsigned*>(ph.amax_out + m), __float_as_uint(rmax));
                    }
                }
            }
            g_op += total;
        }
        g_tile += (uint32_t)ntl;
        if (phx + 1 < p.n_phases) {
            __syncwarp();
            cluster_sync_all();
            if (warp == 0 && lane == 0) {
                const CUtensorMap* nA = phx == 0 ? &mapA1 : &mapA2;
                const CUtensorMap* nWh = phx == 0 ? &mapW1h : &mapW2h;
                const CUtensorMap* nWl = phx == 0 ? &mapW1l : &mapW2l;
                asm volatile("prefetch.tensormap [%0];" ::"l"(nA) : "memory");
                asm volatile("prefetch.tensormap [%0];" ::"l"(nWh) : "memory");
                asm volatile("prefetch.tensormap [%0];" ::"l"(nWl) : "memory");
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

}  // namespace

extern "C" int mi_node_chain(int M, int H, int n_phases, const float* agg, int ld_agg, const float* amax_agg, const void* wb_hi,
                             const void* wb_lo, int ld_wb, const float* bn1, const float* R_, int ld_r, float* an1,
                             float* amax_an1, const void* w2_hi, const void* w2_lo, const float* bn2, const float* h_in,
                             int ld_hin, float* h, int ld_h,
                             const float* ln_g, const float* ln_b, float ln_eps, const void* wpqr_hi, const void* wpqr_lo,
                             const float* cb, int ld_cb, const int* node_graph, float* pqr, int ld_pqr, float* zero_out,
                             int ld_zero, mi_stream_t stream) {
    MI_CHECK_ARG(M >= 0 && (n_phases == 2 || n_phases == 3), "bad sizes");
    MI_CHECK_ARG(H == HMAX, "the fused node chain is built for hidden_dim 512 (four 128-column tiles per cluster)");
    if (M == 0) return MI_OK;
    MI_CHECK_ARG(agg && amax_agg && wb_hi && wb_lo && R_ && an1 && amax_an1 && w2_hi && w2_lo && h && h_in, "null pointer");
    MI_CHECK_ARG(ld_hin % 4 == 0 && mi_host_aligned16(h_in), "operands need 16-byte aligned rows");
    MI_CHECK_ARG(n_phases == 2 || (ln_g && ln_b && wpqr_hi && wpqr_lo && cb && node_graph && pqr), "null pointer (phase 2)");
    MI_CHECK_ARG(ld_agg % 4 == 0 && ld_r % 4 == 0 && ld_h % 4 == 0 && ld_wb % 8 == 0 && (n_phases == 2 || (ld_cb % 4 == 0 && ld_pqr % 4 == 0)) &&
                 mi_host_aligned16(agg) && mi_host_aligned16(R_) && mi_host_aligned16(an1) && mi_host_aligned16(h) &&
                 mi_host_aligned16(wb_hi) && mi_host_aligned16(wb_lo) && (!bn1 || mi_host_aligned16(bn1)) &&
                 (!bn2 || mi_host_aligned16(bn2)) && (n_phases == 2 || (mi_host_aligned16(cb) && mi_host_aligned16(pqr))) &&
                 (!zero_out || (ld_zero % 4 == 0 && mi_host_aligned16(zero_out))),
                 "operands need 16-byte aligned rows");
    int rc = mi_tc_get_encode();
    if (rc != MI_OK) return rc;
    CUtensorMap mA0, mA1, mA2, mW0h, mW0l, mW1h, mW1l, mW2h, mW2l;
    if ((rc = mi_tc_make_map(&mA0, agg, M, H, ld_agg, TM, false)) != MI_OK) return rc;
    if ((rc = mi_tc_make_map(&mA1, an1, M, H, H, TM, false)) != MI_OK) return rc;
    if ((rc = mi_tc_make_map(&mA2, h, M, H, ld_h, TM, false)) != MI_OK) return rc;
    if ((rc = mi_tc_make_map(&mW0h, wb_hi, H, H, ld_wb, TN, true)) != MI_OK) return rc;
    if ((rc = mi_tc_make_map(&mW0l, wb_lo, H, H, ld_wb, TN, true)) != MI_OK) return rc;
    if ((rc = mi_tc_make_map(&mW1h, w2_hi, H, H, H, TN, true)) != MI_OK) return rc;
    if ((rc = mi_tc_make_map(&mW1l, w2_lo, H, H, H, TN, true)) != MI_OK) return rc;
    if (n_phases == 3) {
        if ((rc = mi_tc_make_map(&mW2h, wpqr_hi, 3 * H, H, H, TN, true)) != MI_OK) return rc;
        if ((rc = mi_tc_make_map(&mW2l, wpqr_lo, 3 * H, H, H, TN, true)) != MI_OK) return rc;
    } else {
        mW2h = mW1h;
        mW2l = mW1l;
    }
    Params p = {};
    p.M = M; p.K = H; p.n_phases = n_phases;
    Phase& a = p.ph[0];
    a.n_tiles = 1; a.N = H; a.C = an1; a.ldc = H; a.bias = bn1; a.g1 = R_; a.g1_idx = nullptr; a.g1_ld = ld_r;
    a.act = MI_ACT_SILU; a.amax_out = amax_an1; a.a_amax = amax_agg;
    Phase& b = p.ph[1];
    b.n_tiles = 1; b.N = H; b.C = h; b.ldc = ld_h; b.bias = bn2; b.resid = h_in; b.resid_ld = ld_hin; b.act = MI_ACT_SILU;
    b.a_amax = amax_an1;
    Phase& c = p.ph[2];
    c.n_tiles = 3; c.N = 3 * H; c.C = pqr; c.ldc = ld_pqr; c.g1 = cb; c.g1_idx = node_graph; c.g1_ld = ld_cb; c.act = MI_ACT_NONE;
    c.a_amax = nullptr;
    p.ln_x = h; p.ln_ldx = ld_h; p.ln_g = ln_g; p.ln_b = ln_b; p.ln_eps = ln_eps;
    p.zero_out = n_phases == 3 ? zero_out : nullptr; p.zero_ld = ld_zero;
    static bool attr = false;
    if (!attr) {
        MI_CUDA(cudaFuncSetAttribute(node_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        attr = true;
    }
    const int row_blocks = mi_div_up(M, TM);
    node_chain_kernel<<<row_blocks * CLUSTER, THREADS, SMEM_BYTES, (cudaStream_t)stream>>>(mA0, mA1, mA2, mW0h, mW0l, mW1h, mW1l,
                                                                                          mW2h, mW2l, p);
    MI_CHECK_LAUNCH();
    return MI_OK;
}
