// The node-level chain of a CSPNet layer boundary in ONE launch (inference):
//
//   prologue  agg (fp32, scatter-mean of the layer)  ->  pre-split fp16 operand; agg is zeroed for the next layer's scatter
//   phase 0   an1   = silu(agg W_b^T + R + b_n1)             node_mlp.0 on [LN(h) | agg]; R = LN(h) W_a^T comes from the
//                                                            previous P|Q|R GEMM (cspnet.py:77-82, forward_graph "node path")
//   phase 1   h    += silu(an1 W_n2^T + b_n2)                node_mlp.2 + residual (cspnet.py:82, 91), then the NEXT layer's
//                                                            LayerNorm (cspnet.py:86-88) of the finished rows
//   phase 2   [P'|Q|R] = LN(h) [W_hi; W_hj; W_a]^T + [C_b|0|0]   the next layer's per-node GEMM
//
// The GEMMs are row-wise, so the only dependency between CTAs is across the column tiles of the same 128 rows: a
// thread-block CLUSTER of four CTAs owns one row block (each CTA 128 of the 512 output columns, 384 of the 1536 in
// phase 2) and the phases are separated by cluster barriers.
//
// What makes it fast (the first version of this kernel took 100 us per launch against 78 us for the four separate
// launches it replaced: every CTA re-split the same fp32 A tiles in its main loop and the LayerNorm statistics ran
// serially on six warps; profiles/r2a_breakdown_*.txt):
//   * every A operand is PRE-SPLIT: each CTA's epilogue writes its 128-column slice of the next phase's operand straight
//     from TMEM (lane = row, 32-byte vector stores) as fp16 (hi, 2^11-scaled lo) pairs, so the main loops are pure
//     TMA -> tcgen05.mma pipelines over a six-deep ring, with the W tiles of the next phase prefetched across the barrier;
//   * the power-of-two row scale the fp16 split needs is taken from an a-priori BOUND of the row maximum instead of the
//     maximum itself (which would need another exchange between the four CTAs):
//        |an1| <= |z| <= amax(agg) max_j ||W_b[j]||_1 + amax([P'|Q|R]) + max |b_n1|        (phase 0 -> 1)
//        |LN(h)| <= sqrt(H) max |gamma| + max |beta|                                          (phase 1 -> 2)
//     A bound 2^s above the true maximum costs nothing until s ~ 12: the split keeps 22 bits below the scaled maximum and
//     fp16's subnormal floor (2^-25 relative to 2^15) only then reaches 2^-24 of the row's true maximum;
//   * LayerNorm statistics: every thread owns one row of its CTA's slice in registers; partial (mean, M2) pairs are
//     exchanged through distributed shared memory and merged with Chan's formula (one cluster barrier).
//
// Arithmetic of the products: mi_tc.cu's two-accumulator format (split-precision FP16 x3 on tcgen05, TMEM accumulators).
// Warps: w0 TMA producer, w1 MMA issuer, w2-9 prologue + epilogue.  H = 512 only; other sizes use the separate kernels.
//
// Two forms of the same chain, chosen by the host from the row-block size (mi_node_chain, below):
//   node_chain_kernel     rows on the MMA's M side, a thread of the epilogue owns one ROW (this first half of the file);
//   node_chain_t_kernel   the operands swapped, rows on the N side, a thread owns one FEATURE (second half): faster while the
//                         row blocks are short (<= 64 rows: the strong-scaling regime, 128 crystals per GPU), with the operand
//                         hand-offs between the phases starting on the CTA's own slice.
// The producer / issuer loops of both run on all 32 lanes with the issuing instructions under elect.sync (mi_tc_common.cuh).
#include <cooperative_groups.h>
#include <stdlib.h>

#include "mi_tc_common.cuh"

using namespace mi_tc;

namespace {

constexpr int TM = 128, TN = 128, TK = 32, H = 512, NKB = H / TK;
constexpr int S = 6;                                          // operand ring depth
constexpr int A_H = TM * TK * 2, W_H = TN * TK * 2, OPB = 2 * A_H + 2 * W_H;
constexpr int EPI_WARPS = 8, EPI_WARP0 = 2, THREADS = (EPI_WARP0 + EPI_WARPS) * 32;
constexpr int RING_BYTES = S * OPB, BAR_BYTES = 256;
constexpr int OFF_BAR = RING_BYTES, OFF_PART = OFF_BAR + BAR_BYTES;
constexpr int CLUSTER = 4, PARTS = 2 * CLUSTER;               // partial LayerNorm statistics per row: 4 CTAs x 2 column groups
constexpr int OFF_CST = OFF_PART + TM * PARTS * 8;            // this CTA's column slice of b_n1 | b_n2 | gamma | beta
constexpr int SMEM_BYTES = OFF_CST + 4 * TN * 4;
static_assert(SMEM_BYTES <= 232448, "does not fit the SM");
static_assert((3 * S + 5) * 8 + 8 <= BAR_BYTES, "barrier block too small");
constexpr uint32_t ACC_COLS = 2 * TN, TMEM_COLS = 2 * ACC_COLS;
constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
constexpr uint32_t IDESC_WIDE = (1u << 4) | ((uint32_t)((2 * TN) >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);   // N = [W_hi | W_lo]

struct Params {
    int M, n_phases;
    int rows;                                // rows per cluster (<= TM, multiple of 8): chosen so that the row blocks fill the machine
    int S, slot_bytes, w_off;                // transposed form: operand ring slots, bytes per slot (A pair | W pair), offset of the W pair
    float* agg; int ld_agg; const float* amax_agg; int zero_agg;
    __half* xs_hi; __half* xs_lo;            // [M, H] operand of phase 0 (agg) and of phase 2 (LN(h))
    __half* ys_hi; __half* ys_lo;            // [M, H] operand of phase 1 (an1)
    const float* bn1; const float* R; int ld_r; const float* amax_pqr;
    const float* bounds;                     // device: {max_j ||W_b[j]||_1, max |b_n1|, sqrt(H) max|gamma| + max|beta| of the next LN}
    const float* bn2; const float* h_in; int ld_hin; float* h; int ld_h;
    const float* ln_g; const float* ln_b; float ln_eps;
    const float* cb; int ld_cb; const int* node_graph; float* pqr; int ld_pqr; float* amax_next;
};

#ifdef MI_NODE_TRACE
// Developer instrumentation (scripts/trace_node.py builds a separate library with -DMI_NODE_TRACE; never in the product
// build): per-CTA SM clock stamps of the phases.
constexpr int NT_SLOTS = 32;
__device__ long long g_ntrace[160 * NT_SLOTS];
#define NTRACE(slot) do { g_ntrace[blockIdx.x * NT_SLOTS + (slot)] = clock64(); } while (0)
#else
#define NTRACE(slot) do {} while (0)
#endif

__device__ __forceinline__ void cluster_sync_all() {
    // every thread of the four CTAs: global writes of this phase (generic proxy) -> visible to the peers' loads, including
    // their TMA loads (async proxy)
    asm volatile("fence.proxy.async;" ::: "memory");
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    asm volatile("fence.proxy.async;" ::: "memory");
}

// power-of-two exponent that brings a row with maximum (or bound) `amax` into [2^14, 2^15): the same rule as mi_tc.cu
__device__ __forceinline__ int exp8(float amax) {
    const int ex = (int)((__float_as_uint(amax) >> 23) & 0xff) - 127;
    return max(-100, min(ex - 14, 100));
}
__device__ __forceinline__ float pow2f(int e) { return __uint_as_float((uint32_t)(127 + e) << 23); }
// bound of max |an1[row]| (see the header); one expression, explicit roundings: every CTA and both phases recompute it
__device__ __forceinline__ float an1_bound(float amax_agg, float amax_pqr, float wb_l1, float b_max) {
    return __fadd_rn(__fmaf_rn(amax_agg, wb_l1, amax_pqr), b_max);
}

__device__ __forceinline__ void ldcg8(const float* p, float* v) {
    asm volatile("ld.global.cg.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "l"(p));
}
__device__ __forceinline__ void ldnc8(const float* p, float* v) {
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "l"(p));
}
__device__ __forceinline__ void st8(void* p, const uint32_t* u) {
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(u[0]), "r"(u[1]), "r"(u[2]), "r"(u[3]),
                 "r"(u[4]), "r"(u[5]), "r"(u[6]), "r"(u[7]) : "memory");
}
__device__ __forceinline__ void st8f(float* p, const float* v) {
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]),
                 "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
}

__global__ void __cluster_dims__(CLUSTER, 1, 1) __launch_bounds__(THREADS, 1)
node_chain_kernel(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapY,
                  const __grid_constant__ CUtensorMap mapW0, const __grid_constant__ CUtensorMap mapW1,
                  const __grid_constant__ CUtensorMap mapW2, const Params p) {
    // every map covers an fp16 (hi, lo) PAIR (3-D, planes = hi / lo): one TMA operation per operand and k-block.  The
    // producer is bound by operations (~170 cycles each), not bytes, at these tile sizes.
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* ring = smem;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    uint64_t* w_full = bars;                  // [S] W tiles of the slot landed
    uint64_t* a_full = bars + S;              // [S] A tiles of the slot landed
    uint64_t* op_empty = bars + 2 * S;        // [S] MMAs done with the slot
    uint64_t* acc_full = bars + 3 * S;        // [2]
    uint64_t* acc_empty = acc_full + 2;       // [2]
    uint64_t* stat_bar = acc_full + 4;        // partial LayerNorm statistics of all four CTAs have landed (st.async)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 5);
    float* cst = reinterpret_cast<float*>(smem + OFF_CST);
    float2* part = reinterpret_cast<float2*>(smem + OFF_PART);      // [TM][PARTS] partial (mean, M2) of the LayerNorm rows

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) NTRACE(0);
    const int crank = (int)(blockIdx.x % CLUSTER);               // rank in the cluster = column slice
    const int m0 = (int)(blockIdx.x / CLUSTER) * p.rows;         // the cluster's row block: R <= TM rows (the MMA tile's other rows are not used)

    if (threadIdx.x == 0) {
        if (smem_u32(smem) & 1023u) __trap();
        for (int s = 0; s < S; ++s) {
            mbar_init(&w_full[s], 1);
            mbar_init(&a_full[s], 1);
            mbar_init(&op_empty[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&acc_full[b], 1);
            mbar_init(&acc_empty[b], EPI_WARPS);
        }
        mbar_init(stat_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapX) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapW0) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();                                                   // agg / amax of the preceding per-edge block are complete
    pdl_launch();
    if (threadIdx.x == 0) NTRACE(1);
    const bool tr = threadIdx.x == EPI_WARP0 * 32;                 // the thread that stamps the epilogue side
    (void)tr;

    // Cluster barriers, the same sequence in every thread:  B0 after the prologue, B1 after phase 0 and, with a LayerNorm
    // phase, B2 (LN(h) operand written) before phase 2.
    if (warp == 0) {
        // ===================== TMA producer =====================
        uint32_t g = 0;                                            // k-blocks through the ring so far
        for (int phx = 0; phx < p.n_phases; ++phx) {
            const int ntl = phx == 2 ? 3 : 1;
            const uint32_t total = (uint32_t)(ntl * NKB);
            const uint32_t pre = total < (uint32_t)S ? total : (uint32_t)S;
            const CUtensorMap* mA = phx == 1 ? &mapY : &mapX;
            const CUtensorMap* mW = phx == 0 ? &mapW0 : (phx == 1 ? &mapW1 : &mapW2);
            auto issue_w = [&](uint32_t loc) {
                const uint32_t gi = g + loc;
                const int s = (int)(gi % (uint32_t)S);
                const int tl = (int)(loc / (uint32_t)NKB), kb = (int)(loc % (uint32_t)NKB);
                const int n0 = (crank * ntl + tl) * TN;
                mbar_wait(&op_empty[s], ((gi / (uint32_t)S) & 1) ^ 1);
                uint8_t* st = ring + s * OPB;
                if (elect_one()) {
                    mbar_expect_tx(&w_full[s], 2 * W_H);
                    tma_load_3d(st + 2 * A_H, mW, &w_full[s], kb * TK, n0, 0);      // W_hi | W_lo
                }
                __syncwarp();
            };
            // the weights do not depend on the previous phase: the first ring-full of W tiles crosses the barrier
            // (all lanes run the loops, one elected lane issues: see elect_one() in mi_tc_common.cuh)
            for (uint32_t loc = 0; loc < pre; ++loc) issue_w(loc);
            cluster_sync_all();                                    // B0 / B1 / B2
            for (uint32_t loc = 0; loc < total; ++loc) {
                if (loc >= pre) issue_w(loc);
                const int s = (int)((g + loc) % (uint32_t)S);
                const int kb = (int)(loc % (uint32_t)NKB);
                uint8_t* st = ring + s * OPB;
                if (elect_one()) {
                    mbar_expect_tx(&a_full[s], (uint32_t)(2 * p.rows * TK * 2));
                    tma_load_3d(st, mA, &a_full[s], kb * TK, m0, 0);                // A_hi (rows x 64 B) | A_lo right behind it
                }
                __syncwarp();
            }
            if (phx + 1 < p.n_phases && elect_one()) {
                const CUtensorMap* nA = phx == 0 ? &mapY : &mapX;
                const CUtensorMap* nW = phx == 0 ? &mapW1 : &mapW2;
                asm volatile("prefetch.tensormap [%0];" ::"l"(nA) : "memory");
                asm volatile("prefetch.tensormap [%0];" ::"l"(nW) : "memory");
            }
            __syncwarp();
            g += total;
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        uint32_t g = 0, gt = 0;
        for (int phx = 0; phx < p.n_phases; ++phx) {
            const int ntl = phx == 2 ? 3 : 1;
            cluster_sync_all();                                    // B0 / B1 / B2
            {
                for (int tl = 0; tl < ntl; ++tl, ++gt) {
                    const uint32_t ab = gt & 1;
                    const uint32_t acc = tmem_base + ab * ACC_COLS;
                    mbar_wait(&acc_empty[ab], ((gt >> 1) & 1) ^ 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    for (int kb = 0; kb < NKB; ++kb, ++g) {
                        const int s = (int)(g % (uint32_t)S);
                        const uint32_t par = (g / (uint32_t)S) & 1;
                        mbar_wait(&w_full[s], par);
                        if (kb == 0 && tl == 0 && lane == 0) NTRACE(17 + 3 * phx);
                        mbar_wait(&a_full[s], par);
                        if (kb == 0 && tl == 0 && lane == 0) NTRACE(18 + 3 * phx);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const uint32_t st = smem_u32(ring + s * OPB);
                        const uint64_t d_ahi = umma_desc(st), d_alo = umma_desc(st + (uint32_t)(p.rows * TK * 2));
                        const uint64_t d_whi = umma_desc(st + 2 * A_H), d_wlo = umma_desc(st + 2 * A_H + W_H);
                        // W_hi and W_lo are adjacent tiles of the slot and the main / correction accumulators adjacent TMEM
                        // columns: ONE 256-wide MMA forms a_hi.w_hi and a_hi.w_lo (A_hi is read from shared memory once
                        // instead of twice: the 128-wide tiles of this kernel run at the shared-memory limit), a 128-wide one
                        // adds a_lo.w_hi to the correction
                        (void)d_wlo;
                        if (elect_one()) {
#pragma unroll
                            for (int k = 0; k < TK / 16; ++k) {
                                const uint64_t adv = (uint64_t)((k * 32) >> 4);
                                umma_f16(acc, d_ahi + adv, d_whi + adv, IDESC_WIDE, (kb | k) != 0);
                                umma_f16(acc + TN, d_alo + adv, d_whi + adv, IDESC, 1u);
                            }
                            umma_commit(&op_empty[s]);
                            if (kb == NKB - 1) umma_commit(&acc_full[ab]);
                        }
                        __syncwarp();
                    }
                    if (tl == ntl - 1 && lane == 0) NTRACE(19 + 3 * phx);
                }
            }
        }
    } else {
        // ===================== prologue + epilogue warps (w2..9) =====================
        // Every stage is a SHORT ROLLED LOOP over 16-column groups.  A launch runs each stage once, so the first version's
        // straight-line epilogues (6 000 instructions, 97 KB of code) ran at instruction-fetch speed, 5-8 cycles per warp
        // instruction (profiles/r2_node_trace.md).  The operands gathered from global memory are fetched one group ahead.
        const int t = threadIdx.x - EPI_WARP0 * 32;           // 0..255
        const int q = warp & 3;                               // TMEM lane quarter this warp may access
        const int cg = (warp - EPI_WARP0) >> 2;               // column group: 64 of the tile's 128 columns
        const int rl = q * 32 + lane;                         // row inside the block: the TMEM lane this thread owns
        const int row = m0 + rl;
        const bool ok = rl < p.rows && row < p.M;
        const float wb_l1 = __ldg(p.bounds), b1_max = __ldg(p.bounds + 1);
        const float ln_bound = p.n_phases == 3 ? __ldg(p.bounds + 2) : 1.0f;

        // per-column constants of this CTA's slice -> shared memory (the cluster barriers invalidate L1: a global load per
        // group would pay an L2 round trip every time)
        if (t < TN) {
            cst[t] = __ldg(p.bn1 + crank * TN + t);
            cst[TN + t] = __ldg(p.bn2 + crank * TN + t);
            if (p.n_phases == 3) {
                cst[2 * TN + t] = __ldg(p.ln_g + crank * TN + t);
                cst[3 * TN + t] = __ldg(p.ln_b + crank * TN + t);
            }
        }
        // ---- prologue: this CTA's 128-column slice of agg -> fp16 (hi, lo) pairs, rows scaled from their maxima
#pragma unroll 1
        for (int pg = 0; pg < 4; ++pg) {
            float4 vv[4];
            float am[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int r = m0 + (pg * 4 + u) * 8 + (t >> 5);
                vv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                am[u] = 0.f;
                if (r < p.M && (pg * 4 + u) * 8 + (t >> 5) < p.rows) {
                    vv[u] = __ldcg(reinterpret_cast<const float4*>(p.agg + (long long)r * p.ld_agg + crank * TN + (t & 31) * 4));
                    am[u] = __ldg(p.amax_agg + r);
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int r = m0 + (pg * 4 + u) * 8 + (t >> 5);
                if (r >= p.M || (pg * 4 + u) * 8 + (t >> 5) >= p.rows) continue;
                const float sc = pow2f(-exp8(am[u]));
                uint2 hh, ll;
                split2<0>(vv[u].x * sc, vv[u].y * sc, hh.x, ll.x);
                split2<0>(vv[u].z * sc, vv[u].w * sc, hh.y, ll.y);
                const long long o = (long long)r * H + crank * TN + (t & 31) * 4;
                *reinterpret_cast<uint2*>(p.xs_hi + o) = hh;
                *reinterpret_cast<uint2*>(p.xs_lo + o) = ll;
                if (p.zero_agg)
                    *reinterpret_cast<float4*>(p.agg + (long long)r * p.ld_agg + crank * TN + (t & 31) * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        if (tr) NTRACE(2);
        cluster_sync_all();                                    // B0
        if (tr) NTRACE(3);

        const float am_agg = ok ? __ldg(p.amax_agg + row) : 0.f;
        const float bound1 = ok ? an1_bound(am_agg, __ldcg(p.amax_pqr + row), wb_l1, b1_max) : 0.f;
        const int e_agg = exp8(am_agg), e_an1 = exp8(bound1);
        const int cbase = crank * TN + cg * 64;               // first of this thread's 64 columns (phases 0 and 1)
        const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(cg * 64);
        uint32_t v[16], w[16];
        float rn[16], r[16];

        // ---- phase 0 epilogue: an1 = silu(acc + R + b_n1), written as the pre-split operand of phase 1
        {
            const float rowsc = pow2f(e_agg), osc = pow2f(-e_an1);
            const float* rrow = p.R + (long long)(ok ? row : 0) * p.ld_r + cbase;
            ldcg8(rrow, rn);
            ldcg8(rrow + 8, rn + 8);
            mbar_wait(&acc_full[0], 0);
            if (tr) NTRACE(4);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
            for (int it = 0; it < 4; ++it) {
                tmem_ld16(trow + (uint32_t)(it * 16), v);
                tmem_ld16(trow + (uint32_t)(it * 16) + TN, w);
#pragma unroll
                for (int j = 0; j < 16; ++j) r[j] = rn[j];
                if (it + 1 < 4) {
                    ldcg8(rrow + (it + 1) * 16, rn);
                    ldcg8(rrow + (it + 1) * 16 + 8, rn + 8);
                }
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                uint32_t hi[8], lo[8];
#pragma unroll
                for (int j = 0; j < 16; j += 4) {
                    const float4 b4 = *reinterpret_cast<const float4*>(cst + cg * 64 + it * 16 + j);
                    const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
                    float a[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const float acc = fmaf(__uint_as_float(w[j + u]), LO_UNSCALE, __uint_as_float(v[j + u]));
                        a[u] = silu_fast(rowsc * acc + bb[u] + r[j + u]) * osc;
                    }
                    split2<0>(a[0], a[1], hi[j / 2], lo[j / 2]);
                    split2<0>(a[2], a[3], hi[j / 2 + 1], lo[j / 2 + 1]);
                }
                if (ok) {
                    const long long o = (long long)row * H + cbase + it * 16;
                    st8(p.ys_hi + o, hi);
                    st8(p.ys_lo + o, lo);
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[0]);
        }
        if (tr) NTRACE(5);
        cluster_sync_all();                                    // B1
        if (tr) NTRACE(6);

        // ---- phase 1 epilogue: h = h_in + silu(acc + b_n2); the finished rows stay in TMEM (over their accumulator) for
        // the two LayerNorm passes that follow (next layer's LayerNorm)
        {
            const float rowsc = pow2f(e_an1);
            const uint32_t tb = trow + ACC_COLS;
            const float* hrow = p.h_in + (long long)(ok ? row : 0) * p.ld_hin + cbase;
            ldcg8(hrow, rn);
            ldcg8(hrow + 8, rn + 8);
            mbar_wait(&acc_full[1], 0);
            if (tr) NTRACE(7);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            float sum = 0.f;
#pragma unroll 1
            for (int it = 0; it < 4; ++it) {
                tmem_ld16(tb + (uint32_t)(it * 16), v);
                tmem_ld16(tb + (uint32_t)(it * 16) + TN, w);
#pragma unroll
                for (int j = 0; j < 16; ++j) r[j] = rn[j];
                if (it + 1 < 4) {
                    ldcg8(hrow + (it + 1) * 16, rn);
                    ldcg8(hrow + (it + 1) * 16 + 8, rn + 8);
                }
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int j = 0; j < 16; j += 4) {
                    const float4 b4 = *reinterpret_cast<const float4*>(cst + TN + cg * 64 + it * 16 + j);
                    const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const float acc = fmaf(__uint_as_float(w[j + u]), LO_UNSCALE, __uint_as_float(v[j + u]));
                        r[j + u] = (ok ? r[j + u] : 0.f) + silu_fast(rowsc * acc + bb[u]);
                        sum += r[j + u];
                    }
                }
                if (ok) {
                    float* o = p.h + (long long)row * p.ld_h + cbase + it * 16;
                    st8f(o, r);
                    st8f(o + 8, r + 8);
                }
                if (p.n_phases == 3) tmem_st16(tb + (uint32_t)(it * 16), r);
            }
            if (p.n_phases == 3) {
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                // partial statistics of this thread's 64 columns -> every CTA of the cluster (distributed shared memory)
                const float mp = sum * (1.0f / 64.0f);
                float m2 = 0.f;
#pragma unroll 1
                for (int it = 0; it < 4; ++it) {
                    tmem_ld16(tb + (uint32_t)(it * 16), v);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int j = 0; j < 16; ++j) m2 = fmaf(__uint_as_float(v[j]) - mp, __uint_as_float(v[j]) - mp, m2);
                }
                // st.async: the data lands in the peer's shared memory and counts on the peer's barrier: no cluster barrier,
                // and nothing waits for the global stores of h above
                const uint32_t laddr = smem_u32(part + rl * PARTS + crank * 2 + cg), lbar = smem_u32(stat_bar);
                if (tr) mbar_expect_tx(stat_bar, (uint32_t)(EPI_WARPS * 32 * CLUSTER * 8));
#pragma unroll
                for (int c = 0; c < CLUSTER; ++c) {
                    uint32_t raddr, rbar;
                    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(laddr), "r"(c));
                    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rbar) : "r"(lbar), "r"(c));
                    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f32 [%0], {%1, %2}, [%3];"
                                 ::"r"(raddr), "f"(mp), "f"(m2), "r"(rbar) : "memory");
                }
                if (tr) NTRACE(8);
                mbar_wait(stat_bar, 0);
                if (tr) NTRACE(9);
                float mean = 0.f;
                float2 pp[PARTS];
#pragma unroll
                for (int c = 0; c < PARTS; ++c) {
                    pp[c] = part[rl * PARTS + c];
                    mean += pp[c].x;
                }
                mean *= 1.0f / (float)PARTS;
                float M2 = 0.f;
#pragma unroll
                for (int c = 0; c < PARTS; ++c) M2 += pp[c].y + 64.0f * (pp[c].x - mean) * (pp[c].x - mean);
                const float rstd = 1.0f / sqrtf(M2 * (1.0f / (float)H) + p.ln_eps);
                const float osc = pow2f(-exp8(ln_bound));
#pragma unroll 1
                for (int it = 0; it < 4; ++it) {
                    tmem_ld16(tb + (uint32_t)(it * 16), v);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    uint32_t hi[8], lo[8];
#pragma unroll
                    for (int j = 0; j < 16; j += 4) {
                        const float4 g4 = *reinterpret_cast<const float4*>(cst + 2 * TN + cg * 64 + it * 16 + j);
                        const float4 b4 = *reinterpret_cast<const float4*>(cst + 3 * TN + cg * 64 + it * 16 + j);
                        const float gg[4] = {g4.x, g4.y, g4.z, g4.w}, bb[4] = {b4.x, b4.y, b4.z, b4.w};
                        float a[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) a[u] = ((__uint_as_float(v[j + u]) - mean) * rstd * gg[u] + bb[u]) * osc;
                        split2<0>(a[0], a[1], hi[j / 2], lo[j / 2]);
                        split2<0>(a[2], a[3], hi[j / 2 + 1], lo[j / 2 + 1]);
                    }
                    if (ok) {
                        const long long o = (long long)row * H + cbase + it * 16;
                        st8(p.xs_hi + o, hi);
                        st8(p.xs_lo + o, lo);
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[1]);
            if (p.n_phases == 3) {
                if (tr) NTRACE(10);
                cluster_sync_all();                            // B2
                if (tr) NTRACE(11);
            }
        }

        // ---- phase 2 epilogue: [P'|Q|R] = acc + [C_b|0|0][crystal of the row]; row maxima for the consumers' bounds
        if (p.n_phases == 3) {
            const float rowsc = pow2f(exp8(ln_bound));
            const int gi = ok ? __ldg(p.node_graph + row) : 0;
            float rmax = 0.f;
#pragma unroll 1
            for (int tl = 0; tl < 3; ++tl) {
                const uint32_t gt = 2u + (uint32_t)tl;
                const uint32_t ab = gt & 1;
                const uint32_t tb = trow + ab * ACC_COLS;
                const int ncol = (crank * 3 + tl) * TN + cg * 64;
                const float* crow = p.cb + (long long)gi * p.ld_cb + ncol;     // [C_b|0|0] of the row's crystal
                ldnc8(crow, rn);
                ldnc8(crow + 8, rn + 8);
                mbar_wait(&acc_full[ab], (gt >> 1) & 1);
                if (tr) NTRACE(12 + tl);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
                for (int it = 0; it < 4; ++it) {
                    tmem_ld16(tb + (uint32_t)(it * 16), v);
                    tmem_ld16(tb + (uint32_t)(it * 16) + TN, w);
#pragma unroll
                    for (int j = 0; j < 16; ++j) r[j] = rn[j];
                    if (it + 1 < 4) {
                        ldnc8(crow + (it + 1) * 16, rn);
                        ldnc8(crow + (it + 1) * 16 + 8, rn + 8);
                    }
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const float acc = fmaf(__uint_as_float(w[j]), LO_UNSCALE, __uint_as_float(v[j]));
                        r[j] = rowsc * acc + r[j];
                        rmax = fmaxf(rmax, fabsf(r[j]));
                    }
                    if (ok) {
                        float* o = p.pqr + (long long)row * p.ld_pqr + ncol + it * 16;
                        st8f(o, r);
                        st8f(o + 8, r + 8);
                    }
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc_empty[ab]);
            }
            if (ok && p.amax_next) atomicMax(reinterpret_cast<unsigned*>(p.amax_next + row), __float_as_uint(rmax));
            if (tr) NTRACE(15);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) NTRACE(16);
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// =====================================================================================================================
// The same chain with the operands SWAPPED: D^T = W A^T (features on the M side, the row block's rows on the N side).
//
// In-kernel clock stamps (profiles/r2m_node_pair_experiment.txt) showed what paces the main loops of the kernel above:
// not bytes, not the ring depth — the tensor pipe itself.  A CTA's tile is 128 rows whatever its row block holds (the row
// blocks are 40-88 rows at the benchmark batches: the machine is filled with short blocks), so two thirds of every MMA
// multiply padding, and every tcgen05.mma carries 26-49 cycles of fixed cost.  With W as the M operand (128 features of
// the CTA's column slice) and the row block as the N operand the MMA time follows the rows that exist:
//     acc[:, 0:2R)      = W_hi . [A_hi ; A_lo]^T     ONE N = 2R instruction: main | correction 1 (A_hi and A_lo are adjacent
//                                                    in the slot, exactly as TMA delivers the (hi, lo) pair)
//     acc[:, R:R+R16)  += W_lo . A_hi^T              N = R rounded up to 16 (the surplus columns are never read)
// i.e. (2R + R16) / 4 cycles of tensor pipe per k16 instead of 192.  TMEM then holds the tile transposed — lane = feature,
// column = row — and the epilogues change shape with it: a thread owns ONE feature and walks the rows in groups of eight,
// so every global access of a warp is one contiguous 128-byte row segment (the row-per-lane epilogues above pay one L1
// wavefront per lane), the epilogue work scales with the rows that exist, biases / LayerNorm weights are per-thread
// registers and everything per-row (scales, statistics, crystal index) is a shared-memory broadcast.  LayerNorm
// statistics: warp shuffles over a warp's 32 features, then 16 partial (mean, M2) pairs per row (4 warps x 4 CTAs)
// through distributed shared memory as before.
constexpr int T_RING = 200 * 1024, T_MAX_S = 12, T_PARTS = 16;
constexpr int T_OFF_BAR = T_RING, T_OFF_PART = T_OFF_BAR + 512;
constexpr int T_OFF_ROW = T_OFF_PART + TM * T_PARTS * 8;      // per-row floats: rowsc0 | osc0 | rowsc1 | mean | rstd | (int) crystal
constexpr int T_SMEM = T_OFF_ROW + 6 * TM * 4;
#ifndef MI_NODE_T_WARPS
#define MI_NODE_T_WARPS 16
#endif
constexpr int T_EPI_WARPS = MI_NODE_T_WARPS, T_THREADS = (EPI_WARP0 + T_EPI_WARPS) * 32, T_CG = T_EPI_WARPS / 4;   // epilogue work is per (feature, row group): T_CG groups in flight
static_assert(T_SMEM <= 232448, "does not fit the SM");
static_assert((3 * T_MAX_S + 6) * 8 + 8 <= 512, "barrier block too small");

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr));
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
}
__device__ __forceinline__ float warp_sum(float x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x;
}
// (hi, 2^11-scaled lo) fp16 pair of one value: split2<0> for a single element
__device__ __forceinline__ void split1(float x, __half& hi, __half& lo) {
    hi = __float2half_rn(x);
    lo = __float2half_rn((x - __half2float(hi)) * LO_SCALE);
}

// the cluster barrier in two halves (every thread: arrive, wait, arrive, wait, ...)
__device__ __forceinline__ void cl_arrive() {
    asm volatile("fence.proxy.async;" ::: "memory");
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
}
__device__ __forceinline__ void cl_wait() {
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    asm volatile("fence.proxy.async;" ::: "memory");
}

__global__ void __cluster_dims__(CLUSTER, 1, 1) __launch_bounds__(T_THREADS, 1)
node_chain_t_kernel(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapY,
                    const __grid_constant__ CUtensorMap mapW0, const __grid_constant__ CUtensorMap mapW1,
                    const __grid_constant__ CUtensorMap mapW2, const Params p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* ring = smem;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + T_OFF_BAR);
    uint64_t* w_full = bars;                  // [S] W tiles of the slot landed
    uint64_t* a_full = bars + T_MAX_S;        // [S] A tiles of the slot landed
    uint64_t* op_empty = bars + 2 * T_MAX_S;  // [S] MMAs done with the slot
    uint64_t* acc_full = bars + 3 * T_MAX_S;  // [2]
    uint64_t* acc_empty = acc_full + 2;       // [2]
    uint64_t* stat_bar = acc_full + 4;        // partial LayerNorm statistics of all four CTAs have landed (st.async)
    uint64_t* own_ready = acc_full + 5;       // this CTA's epilogue warps have written their slice of the next phase's operand
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 6);
    float2* part = reinterpret_cast<float2*>(smem + T_OFF_PART);    // [TM][T_PARTS] partial (mean, M2) of the LayerNorm rows
    float* s_rowsc0 = reinterpret_cast<float*>(smem + T_OFF_ROW);   // 2^e of the row's agg operand
    float* s_osc0 = s_rowsc0 + TM;                                  // 2^-e of the row's an1 operand
    float* s_rowsc1 = s_rowsc0 + 2 * TM;                            // 2^e of the row's an1 operand
    float* s_mean = s_rowsc0 + 3 * TM;
    float* s_rstd = s_rowsc0 + 4 * TM;
    int* s_graph = reinterpret_cast<int*>(s_rowsc0 + 5 * TM);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) NTRACE(0);
    const int crank = (int)(blockIdx.x % CLUSTER);               // rank in the cluster = column slice
    const int m0 = (int)(blockIdx.x / CLUSTER) * p.rows;         // the cluster's row block
    const int R = p.rows;                                         // multiple of 8, <= TM
    const uint32_t S = (uint32_t)p.S;

    if (threadIdx.x == 0) {
        if (smem_u32(smem) & 1023u) __trap();
        for (uint32_t s = 0; s < S; ++s) {
            mbar_init(&w_full[s], 1);
            mbar_init(&a_full[s], 1);
            mbar_init(&op_empty[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&acc_full[b], 1);
            mbar_init(&acc_empty[b], T_EPI_WARPS);
        }
        mbar_init(stat_bar, 1);
        mbar_init(own_ready, T_EPI_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapX) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapW0) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();                                                   // agg / amax of the preceding per-edge block are complete
    pdl_launch();
    if (threadIdx.x == 0) NTRACE(1);
    const bool tr = threadIdx.x == EPI_WARP0 * 32;                 // the thread that stamps the epilogue side
    (void)tr;

    // Operand hand-offs between the phases: B0 after the prologue (agg -> xs), B1 after phase 0 (an1 -> ys) and, with a LayerNorm
    // phase, B2 (LN(h) -> xs) before phase 2.  Every CTA writes its 128-column slice of the operand — i.e. 4 of the next
    // phase's 16 k-blocks — and the cluster barrier (release of those global stores to the peers' TMA reads) costs 3-4k cycles.
    // The barrier is therefore split: a CTA starts the next phase on its OWN slice as soon as its own epilogue warps have
    // written it (own_ready: CTA-local), and only the other 12 k-blocks wait for the cluster.  The k-blocks of a phase are
    // walked from the CTA's own slice on: kb = (4 rank + i) mod 16.  Every thread arrives and waits in strict alternation; the
    // MMA and epilogue warps need nothing the barrier orders, so they wait late (when it has long completed).
    if (warp == 0) {
        // ===================== TMA producer =====================
        uint32_t ws = 0, wph = 0, as = 0;                          // ring positions of the W and the A loads (slot, wrap parity)
        for (int phx = 0; phx < p.n_phases; ++phx) {
            const int ntl = phx == 2 ? 3 : 1;
            const uint32_t total = (uint32_t)(ntl * NKB);
            const uint32_t pre = total < S ? total : S;
            const CUtensorMap* mA = phx == 1 ? &mapY : &mapX;
            const CUtensorMap* mW = phx == 0 ? &mapW0 : (phx == 1 ? &mapW1 : &mapW2);
            auto issue_w = [&](uint32_t loc) {
                const int tl = (int)(loc / (uint32_t)NKB), kb = (int)((loc + 4u * (uint32_t)crank) % (uint32_t)NKB);
                const int n0 = (crank * ntl + tl) * TN;
                mbar_wait(&op_empty[ws], wph ^ 1);
                if (elect_one()) {
                    mbar_expect_tx(&w_full[ws], 2 * W_H);
                    tma_load_3d(ring + ws * p.slot_bytes + p.w_off, mW, &w_full[ws], kb * TK, n0, 0);  // W_hi | W_lo
                }
                __syncwarp();
                if (++ws == S) { ws = 0; wph ^= 1; }
            };
            // the weights do not depend on the previous phase: the first ring-full of W tiles crosses the barrier
            // (all lanes run the loops, one elected lane issues: see elect_one())
            cl_arrive();                                           // B0 / B1 / B2: nothing of this warp to release
            for (uint32_t loc = 0; loc < pre; ++loc) issue_w(loc);
            mbar_wait(own_ready, (uint32_t)phx & 1u);              // own slice written (generic -> async proxy fenced by the writers)
            auto issue_a = [&](uint32_t loc) {
                if (loc >= pre) issue_w(loc);
                const int kb = (int)((loc + 4u * (uint32_t)crank) % (uint32_t)NKB);
                if (elect_one()) {
                    mbar_expect_tx(&a_full[as], (uint32_t)(2 * R * TK * 2));
                    tma_load_3d(ring + as * p.slot_bytes, mA, &a_full[as], kb * TK, m0, 0);     // A_hi (R x 64 B) | A_lo right behind it
                }
                __syncwarp();
                if (++as == S) as = 0;
            };
            for (uint32_t loc = 0; loc < 4; ++loc) issue_a(loc);   // own slice
            cl_wait();                                             // the peers' slices
            for (uint32_t loc = 4; loc < total; ++loc) issue_a(loc);
            if (phx + 1 < p.n_phases && elect_one()) {
                const CUtensorMap* nA = phx == 0 ? &mapY : &mapX;
                const CUtensorMap* nW = phx == 0 ? &mapW1 : &mapW2;
                asm volatile("prefetch.tensormap [%0];" ::"l"(nA) : "memory");
                asm volatile("prefetch.tensormap [%0];" ::"l"(nW) : "memory");
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        // kind::f16, D = f32, both operands K-major fp16; M = 128 features, N = rows
        const uint32_t r16 = (uint32_t)((R + 15) / 16 * 16);
        const uint32_t idesc_wide = (1u << 4) | ((uint32_t)((2 * R) >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
        const uint32_t idesc_lo = (1u << 4) | ((r16 >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
        uint32_t gt = 0, s = 0, par = 0;                           // tiles so far; ring position (slot, wrap parity)
        for (int phx = 0; phx < p.n_phases; ++phx) {
            const int ntl = phx == 2 ? 3 : 1;
            cl_arrive();                                           // B0 / B1 / B2 (waited for after the phase's MMAs)
            {
                for (int tl = 0; tl < ntl; ++tl, ++gt) {
                    const uint32_t ab = gt & 1;
                    const uint32_t acc = tmem_base + ab * ACC_COLS;
                    mbar_wait(&acc_empty[ab], ((gt >> 1) & 1) ^ 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    for (int kb = 0; kb < NKB; ++kb) {
                        mbar_wait(&w_full[s], par);
                        if (kb == 0 && tl == 0 && lane == 0) NTRACE(17 + 3 * phx);
                        mbar_wait(&a_full[s], par);
                        if (kb == 0 && tl == 0 && lane == 0) NTRACE(18 + 3 * phx);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const uint32_t st = smem_u32(ring + s * p.slot_bytes);
                        const uint64_t d_rows = umma_desc(st);                    // A_hi rows, then A_lo rows
                        const uint64_t d_whi = umma_desc(st + p.w_off), d_wlo = umma_desc(st + p.w_off + W_H);
                        if (elect_one()) {
#pragma unroll
                            for (int k = 0; k < TK / 16; ++k) {
                                const uint64_t adv = (uint64_t)((k * 32) >> 4);
                                umma_f16(acc, d_whi + adv, d_rows + adv, idesc_wide, (kb | k) != 0);
                                umma_f16(acc + (uint32_t)R, d_wlo + adv, d_rows + adv, idesc_lo, 1u);
                            }
                            umma_commit(&op_empty[s]);
                            if (kb == NKB - 1) umma_commit(&acc_full[ab]);
                        }
                        __syncwarp();
                        if (++s == S) { s = 0; par ^= 1; }
                    }
                    if (tl == ntl - 1 && lane == 0) NTRACE(19 + 3 * phx);
                }
            }
            __syncwarp();
            cl_wait();
        }
    } else {
        // ===================== prologue + epilogue warps (w2..17) =====================
        const int t = threadIdx.x - EPI_WARP0 * 32;           // 0..511
        const int q = warp & 3;                               // TMEM lane quarter this warp may access
        const int cg = (warp - EPI_WARP0) >> 2;               // which of the 8-row groups this warp walks: cg, cg + 4, ...
        const int fl = q * 32 + lane;                         // feature inside the CTA's 128-column slice: the TMEM lane of this thread
        const int nvalid = min(R, p.M - m0);                  // rows of the block that exist (the last block may be short)
        const int ngroups = R >> 3;
        const float wb_l1 = __ldg(p.bounds), b1_max = __ldg(p.bounds + 1);
        const float ln_bound = p.n_phases == 3 ? __ldg(p.bounds + 2) : 1.0f;
        const int fcol = crank * TN + fl;                     // this thread's feature (phases 0 and 1)
        const float b1 = __ldg(p.bn1 + fcol), b2 = __ldg(p.bn2 + fcol);
        const float lg = p.n_phases == 3 ? __ldg(p.ln_g + fcol) : 0.f, lb = p.n_phases == 3 ? __ldg(p.ln_b + fcol) : 0.f;
        const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16);

        // per-row scalars of the block -> shared memory
        if (t < R) {
            const bool okr = t < nvalid;
            const float am_agg = okr ? __ldg(p.amax_agg + m0 + t) : 0.f;
            const float bound1 = okr ? an1_bound(am_agg, __ldcg(p.amax_pqr + m0 + t), wb_l1, b1_max) : 0.f;
            const int e_agg = exp8(am_agg), e_an1 = exp8(bound1);
            s_rowsc0[t] = pow2f(e_agg);
            s_osc0[t] = pow2f(-e_an1);
            s_rowsc1[t] = pow2f(e_an1);
            s_graph[t] = (okr && p.n_phases == 3) ? __ldg(p.node_graph + m0 + t) : 0;
        }
        // ---- prologue: this CTA's 128-column slice of agg -> fp16 (hi, lo) pairs, rows scaled from their maxima
#pragma unroll 1
        for (int pg = 0; pg < TM / (4 * T_EPI_WARPS); ++pg) {
            float4 vv[4];
            float am[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int rl = (pg * 4 + u) * T_EPI_WARPS + (t >> 5);
                vv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                am[u] = 0.f;
                if (rl < nvalid) {
                    vv[u] = __ldcg(reinterpret_cast<const float4*>(p.agg + (long long)(m0 + rl) * p.ld_agg + crank * TN + (t & 31) * 4));
                    am[u] = __ldg(p.amax_agg + m0 + rl);
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int rl = (pg * 4 + u) * T_EPI_WARPS + (t >> 5);
                if (rl >= nvalid) continue;
                const float sc = pow2f(-exp8(am[u]));
                uint2 hh, ll;
                split2<0>(vv[u].x * sc, vv[u].y * sc, hh.x, ll.x);
                split2<0>(vv[u].z * sc, vv[u].w * sc, hh.y, ll.y);
                const long long o = (long long)(m0 + rl) * H + crank * TN + (t & 31) * 4;
                *reinterpret_cast<uint2*>(p.xs_hi + o) = hh;
                *reinterpret_cast<uint2*>(p.xs_lo + o) = ll;
                if (p.zero_agg)
                    *reinterpret_cast<float4*>(p.agg + (long long)(m0 + rl) * p.ld_agg + crank * TN + (t & 31) * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        // hand-off of this CTA's slice of an operand: own producer first (CTA-local), then the cluster
        auto handoff = [&]() {
            asm volatile("fence.proxy.async;" ::: "memory");   // this thread's global stores -> visible to TMA reads
            __syncwarp();
            if (lane == 0) mbar_arrive(own_ready);
            cl_arrive();
        };
        if (tr) NTRACE(2);
        handoff();                                             // B0
        asm volatile("bar.sync 1, %0;" ::"n"(T_EPI_WARPS * 32) : "memory");     // the per-row scalars above
        if (tr) NTRACE(3);

        uint32_t v[8], w[8];
        float gn[8], x[8];

        // ---- phase 0 epilogue: an1 = silu(acc + R + b_n1), written as the pre-split operand of phase 1
        {
            const float* rcol = p.R + fcol;
            auto fetch = [&](int j) {
#pragma unroll
                for (int u = 0; u < 8; ++u) gn[u] = (j * 8 + u < nvalid) ? __ldcg(rcol + (long long)(m0 + j * 8 + u) * p.ld_r) : 0.f;
            };
            if (cg < ngroups) fetch(cg);
            mbar_wait(&acc_full[0], 0);
            if (tr) NTRACE(4);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
            for (int j = cg; j < ngroups; j += T_CG) {
                const int r0 = j * 8;
                tmem_ld8(tlane + (uint32_t)r0, v);
                tmem_ld8(tlane + (uint32_t)(R + r0), w);
#pragma unroll
                for (int u = 0; u < 8; ++u) x[u] = gn[u];
                if (j + T_CG < ngroups) fetch(j + T_CG);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const float acc = fmaf(__uint_as_float(w[u]), LO_UNSCALE, __uint_as_float(v[u]));
                    const float a = silu_fast(s_rowsc0[r0 + u] * acc + b1 + x[u]) * s_osc0[r0 + u];
                    __half hi, lo;
                    split1(a, hi, lo);
                    if (r0 + u < nvalid) {
                        const long long o = (long long)(m0 + r0 + u) * H + fcol;
                        p.ys_hi[o] = hi;
                        p.ys_lo[o] = lo;
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[0]);
        }
        if (tr) NTRACE(5);
        cl_wait();                                             // (B0, long complete)
        handoff();                                             // B1
        if (tr) NTRACE(6);

        // ---- phase 1 epilogue: h = h_in + silu(acc + b_n2); the finished values stay in TMEM (over their accumulator) for
        // the LayerNorm pass that follows (next layer's LayerNorm)
        {
            const uint32_t tb = tlane + ACC_COLS;
            const float* hcol = p.h_in + fcol;
            float* ocol = p.h + fcol;
            auto fetch = [&](int j) {
#pragma unroll
                for (int u = 0; u < 8; ++u) gn[u] = (j * 8 + u < nvalid) ? __ldcg(hcol + (long long)(m0 + j * 8 + u) * p.ld_hin) : 0.f;
            };
            if (cg < ngroups) fetch(cg);
            if (p.n_phases == 3 && tr) mbar_expect_tx(stat_bar, (uint32_t)(R * T_PARTS * 8));
            mbar_wait(&acc_full[1], 0);
            if (tr) NTRACE(7);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
            for (int j = cg; j < ngroups; j += T_CG) {
                const int r0 = j * 8;
                tmem_ld8(tb + (uint32_t)r0, v);
                tmem_ld8(tb + (uint32_t)(R + r0), w);
#pragma unroll
                for (int u = 0; u < 8; ++u) x[u] = gn[u];
                if (j + T_CG < ngroups) fetch(j + T_CG);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const float acc = fmaf(__uint_as_float(w[u]), LO_UNSCALE, __uint_as_float(v[u]));
                    x[u] += silu_fast(s_rowsc1[r0 + u] * acc + b2);
                    if (r0 + u < nvalid) ocol[(long long)(m0 + r0 + u) * p.ld_h] = x[u];
                }
                if (p.n_phases == 3) {
                    tmem_st8(tb + (uint32_t)r0, x);
                    // statistics of the eight rows over this warp's 32 features; lane u carries row u's pair to the four CTAs
                    float smp = 0.f, sm2 = 0.f;
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const float mp = warp_sum(x[u]) * (1.0f / 32.0f);
                        const float d = x[u] - mp;
                        const float m2 = warp_sum(d * d);
                        if (lane == u) { smp = mp; sm2 = m2; }
                    }
                    if (lane < 8) {
                        const uint32_t laddr = smem_u32(part + (r0 + lane) * T_PARTS + crank * 4 + q), lbar = smem_u32(stat_bar);
#pragma unroll
                        for (int c = 0; c < CLUSTER; ++c) {
                            uint32_t raddr, rbar;
                            asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(laddr), "r"(c));
                            asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rbar) : "r"(lbar), "r"(c));
                            asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f32 [%0], {%1, %2}, [%3];"
                                         ::"r"(raddr), "f"(smp), "f"(sm2), "r"(rbar) : "memory");
                        }
                    }
                }
            }
            if (p.n_phases == 3) {
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                if (tr) NTRACE(8);
                mbar_wait(stat_bar, 0);
                if (tr) NTRACE(9);
                if (t < R) {                                      // Chan's merge of the row's 16 partial pairs (32 features each)
                    float2 pp[T_PARTS];
                    float mean = 0.f;
#pragma unroll
                    for (int c = 0; c < T_PARTS; ++c) {
                        pp[c] = part[t * T_PARTS + c];
                        mean += pp[c].x;
                    }
                    mean *= 1.0f / (float)T_PARTS;
                    float M2 = 0.f;
#pragma unroll
                    for (int c = 0; c < T_PARTS; ++c) M2 += pp[c].y + 32.0f * (pp[c].x - mean) * (pp[c].x - mean);
                    s_mean[t] = mean;
                    s_rstd[t] = 1.0f / sqrtf(M2 * (1.0f / (float)H) + p.ln_eps);
                }
                asm volatile("bar.sync 1, %0;" ::"n"(T_EPI_WARPS * 32) : "memory");
                const float osc = pow2f(-exp8(ln_bound));
#pragma unroll 1
                for (int j = cg; j < ngroups; j += T_CG) {
                    const int r0 = j * 8;
                    tmem_ld8(tb + (uint32_t)r0, v);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const float a = ((__uint_as_float(v[u]) - s_mean[r0 + u]) * s_rstd[r0 + u] * lg + lb) * osc;
                        __half hi, lo;
                        split1(a, hi, lo);
                        if (r0 + u < nvalid) {
                            const long long o = (long long)(m0 + r0 + u) * H + fcol;
                            p.xs_hi[o] = hi;
                            p.xs_lo[o] = lo;
                        }
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[1]);
            cl_wait();                                         // (B1)
            if (p.n_phases == 3) {
                if (tr) NTRACE(10);
                handoff();                                     // B2
                if (tr) NTRACE(11);
            }
        }

        // ---- phase 2 epilogue: [P'|Q|R] = acc + [C_b|0|0][crystal of the row]; row maxima for the consumers' bounds
        if (p.n_phases == 3) {
            const float rowsc = pow2f(exp8(ln_bound));
#pragma unroll 1
            for (int tl = 0; tl < 3; ++tl) {
                const uint32_t gt = 2u + (uint32_t)tl;
                const uint32_t ab = gt & 1;
                const uint32_t tb = tlane + ab * ACC_COLS;
                const int ncol = (crank * 3 + tl) * TN + fl;
                const float* ccol = p.cb + ncol;                               // [C_b|0|0], row = the crystal of the node
                auto fetch = [&](int j) {
#pragma unroll
                    for (int u = 0; u < 8; ++u)
                        gn[u] = (j * 8 + u < nvalid) ? __ldg(ccol + (long long)s_graph[j * 8 + u] * p.ld_cb) : 0.f;
                };
                if (cg < ngroups) fetch(cg);
                mbar_wait(&acc_full[ab], (gt >> 1) & 1);
                if (tr) NTRACE(12 + tl);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
                for (int j = cg; j < ngroups; j += T_CG) {
                    const int r0 = j * 8;
                    tmem_ld8(tb + (uint32_t)r0, v);
                    tmem_ld8(tb + (uint32_t)(R + r0), w);
#pragma unroll
                    for (int u = 0; u < 8; ++u) x[u] = gn[u];
                    if (j + T_CG < ngroups) fetch(j + T_CG);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    uint32_t mx = 0;
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const float acc = fmaf(__uint_as_float(w[u]), LO_UNSCALE, __uint_as_float(v[u]));
                        x[u] = rowsc * acc + x[u];
                        if (r0 + u < nvalid) p.pqr[(long long)(m0 + r0 + u) * p.ld_pqr + ncol] = x[u];
                        const uint32_t m = __reduce_max_sync(0xffffffffu, __float_as_uint(fabsf(x[u])));
                        if (lane == u) mx = m;
                    }
                    if (lane < 8 && r0 + lane < nvalid && p.amax_next)
                        atomicMax(reinterpret_cast<unsigned*>(p.amax_next + m0 + r0 + lane), mx);
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc_empty[ab]);
            }
            if (tr) NTRACE(15);
            cl_wait();                                         // (B2)
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) NTRACE(16);
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

}  // namespace

#ifdef MI_NODE_TRACE
extern "C" int mi_node_trace_read(long long* out, int n) {
    cudaDeviceSynchronize();
    return (int)cudaMemcpyFromSymbol(out, g_ntrace, sizeof(long long) * (size_t)n);
}
#endif

extern "C" int mi_node_chain(int M, int Hdim, int n_phases, float* agg, int ld_agg, const float* amax_agg, int zero_agg,
                             void* xs_hi, void* xs_lo, void* ys_hi, void* ys_lo, const void* wb_hi, const void* wb_lo, int ld_wb,
                             const float* bn1, const float* R_, int ld_r, const float* amax_pqr, const float* bounds,
                             const void* w2_hi, const void* w2_lo, const float* bn2, const float* h_in, int ld_hin, float* h,
                             int ld_h, const float* ln_g, const float* ln_b, float ln_eps, const void* wpqr_hi,
                             const void* wpqr_lo, const float* cb, int ld_cb, const int* node_graph, float* pqr, int ld_pqr,
                             float* amax_next, mi_stream_t stream) {
    MI_CHECK_ARG(M >= 0 && (n_phases == 2 || n_phases == 3), "bad sizes");
    MI_CHECK_ARG(Hdim == H, "the fused node chain is built for hidden_dim 512 (four 128-column tiles per cluster)");
    if (M == 0) return MI_OK;
    MI_CHECK_ARG(agg && amax_agg && xs_hi && xs_lo && ys_hi && ys_lo && wb_hi && wb_lo && bn1 && R_ && amax_pqr && bounds &&
                 w2_hi && w2_lo && bn2 && h && h_in, "null pointer");
    MI_CHECK_ARG(n_phases == 2 || (ln_g && ln_b && wpqr_hi && wpqr_lo && cb && node_graph && pqr), "null pointer (phase 2)");
    // 32-byte vector accesses of the row-per-lane epilogues, 16-byte TMA rows
    auto al32 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 31) == 0; };
    MI_CHECK_ARG(ld_agg % 4 == 0 && ld_r % 8 == 0 && ld_h % 8 == 0 && ld_hin % 8 == 0 && ld_wb % 8 == 0 &&
                 (n_phases == 2 || (ld_cb % 8 == 0 && ld_pqr % 8 == 0)) && mi_host_aligned16(agg) && al32(R_) && al32(h) &&
                 al32(h_in) && al32(xs_hi) && al32(xs_lo) && al32(ys_hi) && al32(ys_lo) && mi_host_aligned16(wb_hi) &&
                 mi_host_aligned16(wb_lo) && mi_host_aligned16(bn1) && mi_host_aligned16(bn2) &&
                 (n_phases == 2 || (al32(cb) && al32(pqr) && mi_host_aligned16(ln_g) && mi_host_aligned16(ln_b))),
                 "operands need 32-byte aligned rows");
    int rc = mi_tc_get_encode();
    if (rc != MI_OK) return rc;
    CUtensorMap mX, mY, mW0, mW1, mW2;
    if ((rc = mi_tc_make_map_pair(&mW0, wb_hi, wb_lo, H, H, ld_wb, TN)) != MI_OK) return rc;
    if ((rc = mi_tc_make_map_pair(&mW1, w2_hi, w2_lo, H, H, H, TN)) != MI_OK) return rc;
    if (n_phases == 3) {
        if ((rc = mi_tc_make_map_pair(&mW2, wpqr_hi, wpqr_lo, 3 * H, H, H, TN)) != MI_OK) return rc;
    } else {
        mW2 = mW1;
    }
    Params p = {};
    p.M = M; p.n_phases = n_phases;
    p.agg = agg; p.ld_agg = ld_agg; p.amax_agg = amax_agg; p.zero_agg = zero_agg;
    p.xs_hi = (__half*)xs_hi; p.xs_lo = (__half*)xs_lo; p.ys_hi = (__half*)ys_hi; p.ys_lo = (__half*)ys_lo;
    p.bn1 = bn1; p.R = R_; p.ld_r = ld_r; p.amax_pqr = amax_pqr; p.bounds = bounds;
    p.bn2 = bn2; p.h_in = h_in; p.ld_hin = ld_hin; p.h = h; p.ld_h = ld_h;
    p.ln_g = ln_g; p.ln_b = ln_b; p.ln_eps = ln_eps;
    p.cb = cb; p.ld_cb = ld_cb; p.node_graph = node_graph; p.pqr = pqr; p.ld_pqr = ld_pqr; p.amax_next = amax_next;
    // Row blocks: the stages of these kernels are latency-bound per CTA (epilogue traffic, operand ingest), so the blocks are
    // made as short as the machine allows: as many clusters as can be co-resident, 8-row granularity.
    static int max_clusters[2] = {0, 0};
    for (int f = 0; f < 2; ++f) {
        if (max_clusters[f]) continue;
        const void* kf = f ? (const void*)node_chain_t_kernel : (const void*)node_chain_kernel;
        MI_CUDA(cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, f ? T_SMEM : SMEM_BYTES));
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(CLUSTER * 64);
        cfg.blockDim = dim3(f ? T_THREADS : THREADS);
        cfg.dynamicSmemBytes = f ? T_SMEM : SMEM_BYTES;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = CLUSTER; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        int nc = 0;
        if (cudaOccupancyMaxActiveClusters(&nc, kf, &cfg) != cudaSuccess || nc <= 0) {
            cudaGetLastError();
            nc = 21;
        }
        max_clusters[f] = nc;
    }
    auto rows_for = [&](int f) {
        int r = (mi_div_up(M, max_clusters[f]) + 7) / 8 * 8;
        const char* force_r = getenv("MI_NODE_ROWS");
        if (force_r) r = atoi(force_r) / 8 * 8;
        return r > TM ? TM : (r < 8 ? 8 : r);
    };
    // Which form: the transposed one (rows on the N side: MMA time and epilogue work follow the rows that exist) while the row
    // blocks are short, the row-per-lane one above for fuller blocks (measured: 37.2 against 38.3 us per launch at 48 rows,
    // 47.7 against 46.7 at 88, 198 against 183 at 128).  MI_NODE_T=0 / 1 forces one.
    static int force_t = -2;
    if (force_t == -2) {
        const char* e = getenv("MI_NODE_T");
        force_t = !e ? -1 : (e[0] == '0' ? 0 : 1);
    }
    const int use_t = force_t >= 0 ? force_t : (rows_for(1) <= 64 ? 1 : 0);
    const int smem_bytes = use_t ? T_SMEM : SMEM_BYTES, threads = use_t ? T_THREADS : THREADS;
    const int R = rows_for(use_t);
    p.rows = R;
    // transposed form: ring slots packed to the row block (A pair: 2 x R x 64 B, W pair: 16 KB)
    p.w_off = (2 * R * TK * 2 + 1023) / 1024 * 1024;
    p.slot_bytes = p.w_off + 2 * W_H;
    p.S = T_RING / p.slot_bytes;
    if (p.S > T_MAX_S) p.S = T_MAX_S;
    if ((rc = mi_tc_make_map_pair(&mX, xs_hi, xs_lo, M, H, H, R)) != MI_OK) return rc;
    if ((rc = mi_tc_make_map_pair(&mY, ys_hi, ys_lo, M, H, H, R)) != MI_OK) return rc;
    const int row_blocks = mi_div_up(M, R);
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3((unsigned)(row_blocks * CLUSTER));
    lc.blockDim = dim3(threads);
    lc.dynamicSmemBytes = smem_bytes;
    lc.stream = (cudaStream_t)stream;
    cudaLaunchAttribute la[1];
    la[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    la[0].val.programmaticStreamSerializationAllowed = 1;
    lc.attrs = la;
    lc.numAttrs = mi_pdl_enabled() ? 1 : 0;
    if (use_t) {
        MI_CUDA(cudaLaunchKernelEx(&lc, node_chain_t_kernel, mX, mY, mW0, mW1, mW2, p));
    } else {
        MI_CUDA(cudaLaunchKernelEx(&lc, node_chain_kernel, mX, mY, mW0, mW1, mW2, p));
    }
    MI_CHECK_LAUNCH();
    return MI_OK;
}
