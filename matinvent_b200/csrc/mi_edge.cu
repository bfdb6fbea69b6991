// The two per-edge blocks of a CSPNet layer (inference) on CTA PAIRS: tcgen05.mma.cta_group::2, 256 x 256 tiles.
//
//   block 1   a1  = silu(Phi W_F^T + P'[src] + Q[dst])         first edge linear, per-edge part (cspnet.py:59-72)
//   block 2   agg = mean_{e in row segment} silu(a1 W_2^T + b_2)   second edge linear + scatter-mean (cspnet.py:73-79)
//
// Same arithmetic as mi_tc.cu's merged format (split-precision FP16 x3 into one TMEM accumulator, operands pre-scaled by
// powers of two); what changes is where the bytes go.  The single-CTA kernel is bound by what an SM can take in and read
// back from shared memory (48 KB of operands per 32-wide k-block of a 128x256 tile, profiles/r1_tc_trace.md).  Here
//   * BOTH operands arrive pre-split by TMA (Phi from mi_edge_fourier, a1 from block 1's epilogue): no split warps, no raw
//     fp32 ring, no conversion traffic through shared memory;
//   * a CTA pair shares W: each SM stages 128 of the tile's 256 weight rows and its own 128 rows of A, 32 KB per k-block
//     instead of 48, and the tensor core reads each staged byte once for 256 output columns;
//   * block 1's epilogue goes straight from TMEM to global memory in the row-per-lane layout tcgen05.ld delivers (32-byte
//     vector loads of the gathered rows, 32-byte vector stores of fp16 (hi, lo) pairs): no transpose through shared memory,
//     no row-maximum atomics.  The power-of-two row scale of a1 comes from an a-priori bound of the row maximum,
//        |a1[e]| <= |z1[e]| <= sqrt(3F) max_j ||W_F[j]||_2 + amax([P'|Q|R][src]) + amax([P'|Q|R][dst])
//     (every sin/cos pair of Phi has unit norm), stored as block 2's `a_bound` — see mi_node.cu for why a bound is enough;
//   * block 2 keeps mi_tc.cu's fused scatter-mean epilogue (per-warp transpose buffer, one 128-byte reduction per segment
//     and 32-column chunk).
// Pipeline: per CTA w0 = TMA producer (both CTAs load their halves, completing on the LEADER's barrier), w1 = MMA issuer
// (leader CTA only; commits are multicast to both CTAs), w2-9 = epilogue; TMEM holds two 256-column accumulators per CTA so
// the epilogue of a tile overlaps the main loop of the next.  Persistent: one pair per two SMs.
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>

#include "mi_tc_common.cuh"

using namespace mi_tc;

namespace {

constexpr int TM = 128;                  // rows per CTA; the pair's tile is 2 TM x TNP
constexpr int TNP = 256;                 // columns of the pair's tile: each CTA stages TNP / 2 weight rows
constexpr int TK = 32;
constexpr int T_H = TM * TK * 2;         // one fp16 tile: 128 rows x 64 bytes (SWIZZLE_64B)
constexpr int OPB = 4 * T_H;             // ring slot: A_hi | A_lo | W_hi | W_lo
constexpr int EPI_WARP0 = 2;              // w0 TMA producer, w1 MMA issuer, then the epilogue warps
constexpr int EP = 34;                   // floats per transpose-buffer row
constexpr int BAR_BYTES = 256;
constexpr uint32_t ACC_COLS = TNP, TMEM_COLS = 2 * ACC_COLS;
// kind::f16, D = f32, A = B = f16, K-major, N = 256, M = 256 (the pair)
constexpr uint32_t IDESC2 = (1u << 4) | ((uint32_t)(TNP >> 3) << 17) | ((uint32_t)((2 * TM) >> 4) << 24);

template <int MODE>
struct ECfg {
    // Epilogue warps.  Block 1's epilogue is latency-bound per warp (TMEM read -> gathers -> SiLU -> split -> staging ->
    // TMA store); with the MMAs pacing the steady state what it costs is the exposed tail after each CTA's last tile
    // (18k of 98k cycles with 8 warps): it works in 16-column groups (112 registers) so that 16 warps fit.  Block 2
    // (transposes + segment sums, 32-column chunks) keeps 8: with 16 its ring would shrink to 4 slots (measured slower).
    static constexpr int EPI_WARPS = MODE == 0 ? 16 : 8;
    static constexpr int THREADS = (EPI_WARP0 + EPI_WARPS) * 32;
    static constexpr int S = MODE == 0 ? 6 : 5;                                   // ring depth
    // block 2: per-warp transpose buffers; block 1: per-warp staging of a 32 x 16 (hi, lo) fp16 output group for TMA stores
    static constexpr int EBUF_BYTES = MODE == 0 ? EPI_WARPS * 2048 : EPI_WARPS * 32 * EP * 4;
    static constexpr int RING_BYTES = S * OPB;
    static constexpr int SMEM_BYTES = RING_BYTES + EBUF_BYTES + BAR_BYTES;
    static_assert(SMEM_BYTES <= 232448, "does not fit the SM");
    static_assert((2 * S + 4) * 8 + 8 <= BAR_BYTES, "barrier block too small");
};

struct EdgeParams {
    int M, N, K;
    float alpha;
    const float* col_scale;
    // block 1
    const float* P; const float* Q; int ld_pq; const int* src; const int* dst; const float* amax_pq; const float* wf_bound;
    __half* out_hi; __half* out_lo; int ld_out; float* bound_out;
    // block 2
    const float* a_bound; const float* bias;
    float* scat_out; int scat_ld; const int* scat_idx; const float* scat_w; float* scat_amax;
};

#ifdef MI_EDGE_TRACE
// Developer instrumentation (scripts/trace_edge.py builds a separate library with -DMI_EDGE_TRACE; never in the product
// build): per-CTA, per-tile SM clock stamps of the pipeline roles.
constexpr int ET_TILES = 6, ET_SLOTS = 8;
__device__ long long g_etrace[160 * ET_TILES * ET_SLOTS];
#define ETRACE(tl, slot) do { if ((tl) < ET_TILES) g_etrace[(blockIdx.x * ET_TILES + (tl)) * ET_SLOTS + (slot)] = clock64(); } while (0)
#else
#define ETRACE(tl, slot) do {} while (0)
#endif

__device__ __forceinline__ uint32_t mapa_rank(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_expect_tx_cluster(uint32_t caddr, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cluster.b64 _, [%0], %1;" ::"r"(caddr), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t caddr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(caddr) : "memory");
}
// TMA load of this CTA's tile whose completion is counted on a barrier of the pair's leader CTA
__device__ __forceinline__ void tma_load_2d_pair(void* dst, const CUtensorMap* map, uint32_t bar_caddr, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(bar_caddr), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on the barrier at this shared-memory offset in BOTH CTAs of the pair once the MMAs issued so far have completed
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
// TMA store of a [32 rows x 16 columns] fp16 tile (SWIZZLE_32B in shared memory); rows / columns past the tensor are clipped
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void cluster_sync_pair() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ int exp8(float amax) {
    const int ex = (int)((__float_as_uint(amax) >> 23) & 0xff) - 127;
    return max(-100, min(ex - 14, 100));
}
__device__ __forceinline__ float pow2f(int e) { return __uint_as_float((uint32_t)(127 + e) << 23); }
__device__ __forceinline__ void ldnc8(const float* p, float* v) {
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "l"(p));
}
// Segment sums of one 32x32 chunk held in a warp's transpose buffer (row-major, pitch EP): lane = column, the rows of
// every segment are walked in order and one 128-byte reduction per (segment, chunk) goes to the zeroed destination
// (the same routine as mi_tc.cu's: a segment meets at most two 32-row windows, so the sum does not depend on order).
__device__ __forceinline__ void scatter_chunk(const float* ebuf, uint32_t starts, int my_seg, float my_w,
                                              float* __restrict__ out, int ld, int nb, int lane) {
    uint32_t m = starts;
    while (m) {                                                  // warp-uniform
        const int a = __ffs(m) - 1;
        m &= m - 1;
        const int b = m ? __ffs(m) - 1 : 32;
        const int sg = __shfl_sync(0xffffffffu, my_seg, a);
        const float wg = __shfl_sync(0xffffffffu, my_w, a);
        if (sg < 0) break;                                       // rows past M
        float s0 = 0.f, s1 = 0.f;
        int rr = a;
        for (; rr + 2 <= b; rr += 2) {
            s0 += ebuf[rr * EP + lane];
            s1 += ebuf[(rr + 1) * EP + lane];
        }
        if (rr < b) s0 += ebuf[rr * EP + lane];
        atomicAdd(out + (long long)sg * ld + nb + lane, (s0 + s1) * wg);
    }
}

// MODE 0: block 1 (gathers + SiLU, pre-split output).  MODE 1: block 2 (bias + SiLU + scatter-mean).
template <int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__((ECfg<MODE>::THREADS), 1)
edge_pair_kernel(const __grid_constant__ CUtensorMap mapAh, const __grid_constant__ CUtensorMap mapAl,
                 const __grid_constant__ CUtensorMap mapWh, const __grid_constant__ CUtensorMap mapWl,
                 const __grid_constant__ CUtensorMap mapOh, const __grid_constant__ CUtensorMap mapOl, const EdgeParams p) {
    using C = ECfg<MODE>;
    constexpr int S = C::S, EPI_WARPS = C::EPI_WARPS;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* ring = smem;
    float* ebuf_all = reinterpret_cast<float*>(smem + C::RING_BYTES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::RING_BYTES + C::EBUF_BYTES);
    uint64_t* full = bars;                    // [S] leader's: both CTAs' tiles of the slot have landed
    uint64_t* empty = bars + S;               // [S] every CTA's own: the MMAs are done with the slot
    uint64_t* acc_full = bars + 2 * S;        // [2] every CTA's own: the accumulators of a tile are complete
    uint64_t* acc_empty = acc_full + 2;       // [2] leader's: both CTAs' epilogue warps have drained the buffer
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = blockIdx.x & 1;                         // rank in the pair (cluster dims 2x1x1 over a 1-D grid)
    const int pair = (int)(blockIdx.x >> 1), npairs = (int)(gridDim.x >> 1);
    const int nkb = (p.K + TK - 1) / TK;
    const int tiles_n = p.N / TNP;
    const int tiles_m = (p.M + 2 * TM - 1) / (2 * TM);
    const int num_tiles = tiles_m * tiles_n;
    const int my_tiles = (num_tiles - pair + npairs - 1) / npairs;
    const uint32_t total = (uint32_t)my_tiles * (uint32_t)nkb;
    auto tile_origin = [&](int tl, int& m0, int& n0) {            // n fastest: the pairs sharing an A row block run together
        const int tile = pair + tl * npairs;
        m0 = (tile / tiles_n) * (2 * TM);
        n0 = (tile % tiles_n) * TNP;
    };

    if (threadIdx.x == 0) {
        if (smem_u32(smem) & 1023u) __trap();
        for (int s = 0; s < S; ++s) {
            mbar_init(&full[s], 2);                               // one arrive.expect_tx per CTA of the pair
            mbar_init(&empty[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&acc_full[b], 1);
            mbar_init(&acc_empty[b], 2 * EPI_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapAh) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapAl) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapWh) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapWl) : "memory");
        if (MODE == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&mapOh) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&mapOl) : "memory");
        }
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_pair();                                          // the peer's barriers are initialised before anyone arrives on them
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();                                                   // the predecessor's output (A operand, P|Q rows) is complete
    pdl_launch();

    if (warp == 0) {
        // ===================== TMA producer (both CTAs) =====================
        // All lanes run the loop, one elected lane issues (mi_tc_common.cuh, elect_one(): inside an `if (lane == 0)` region
        // every TMA / MMA instruction is wrapped in a R2UR.BROADCAST loop, ~170 cycles per TMA operation, ~90 per MMA)
        for (uint32_t it = 0; it < total; ++it) {
            const int tl = (int)(it / (uint32_t)nkb), kb = (int)(it % (uint32_t)nkb);
            int m0, n0;
            tile_origin(tl, m0, n0);
            const int s = (int)(it % (uint32_t)S);
            mbar_wait(&empty[s], ((it / (uint32_t)S) & 1) ^ 1);
            if (lane == 0) {
                if (kb == 0) ETRACE(tl, 0);
                if (kb == nkb - 1) ETRACE(tl, 1);
            }
            uint8_t* st = ring + s * OPB;
            const uint32_t fb = mapa_rank(smem_u32(&full[s]), 0);
            const int kc = kb * TK, ma = m0 + (int)rank * TM, nw = n0 + (int)rank * (TNP / 2);
            if (elect_one()) {
                mbar_expect_tx_cluster(fb, OPB);
                tma_load_2d_pair(st, &mapAh, fb, kc, ma);
                tma_load_2d_pair(st + T_H, &mapAl, fb, kc, ma);
                tma_load_2d_pair(st + 2 * T_H, &mapWh, fb, kc, nw);
                tma_load_2d_pair(st + 3 * T_H, &mapWl, fb, kc, nw);
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA) =====================
        if (rank == 0) {                                          // all lanes run the loop, one elected lane issues
            uint32_t it = 0;
            for (int tl = 0; tl < my_tiles; ++tl) {
                const uint32_t ab = (uint32_t)tl & 1;
                const uint32_t acc = tmem_base + ab * ACC_COLS;
                if (lane == 0) ETRACE(tl, 2);
                mbar_wait(&acc_empty[ab], (((uint32_t)tl >> 1) & 1) ^ 1);
                if (lane == 0) ETRACE(tl, 3);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int s = (int)(it % (uint32_t)S);
                    mbar_wait(&full[s], (it / (uint32_t)S) & 1);
                    if (kb == 0 && lane == 0) ETRACE(tl, 4);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t st = smem_u32(ring + s * OPB);
                    const uint64_t d_ahi = umma_desc(st), d_alo = umma_desc(st + T_H);
                    const uint64_t d_whi = umma_desc(st + 2 * T_H), d_wlo = umma_desc(st + 3 * T_H);
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < TK / 16; ++k) {
                            const uint64_t adv = (uint64_t)((k * 32) >> 4);
                            umma_f16_pair(acc, d_ahi + adv, d_whi + adv, IDESC2, (kb | k) != 0);
                            umma_f16_pair(acc, d_alo + adv, d_whi + adv, IDESC2, 1u);
                            umma_f16_pair(acc, d_ahi + adv, d_wlo + adv, IDESC2, 1u);
                        }
                        umma_commit_pair(&empty[s]);
                        if (kb == nkb - 1) umma_commit_pair(&acc_full[ab]);
                    }
                    __syncwarp();
                }
                if (lane == 0) ETRACE(tl, 5);
            }
        }
    } else {
        // ===================== epilogue warps (w2..9), overlapped with the next tile's main loop =====================
        const int q = warp & 3;                                   // TMEM lane quarter this warp may access
        const int cg = (warp - EPI_WARP0) >> 2;                   // column group of the tile: 256 / (EPI_WARPS / 4) columns
        float* ebuf = ebuf_all + (warp - EPI_WARP0) * (32 * EP);
        (void)ebuf;
        for (int tl = 0; tl < my_tiles; ++tl) {
            int m0, n0;
            tile_origin(tl, m0, n0);
            const uint32_t ab = (uint32_t)tl & 1;
            const int mrow = m0 + (int)rank * TM + q * 32 + lane;       // the row this lane owns in TMEM
            const bool ok = mrow < p.M;
            constexpr int WCOLS = TNP / (EPI_WARPS / 4);          // columns per warp: 64 (block 1), 128 (block 2)
            const uint32_t tbase = tmem_base + ab * ACC_COLS + ((uint32_t)(q * 32) << 16) + (uint32_t)(cg * WCOLS);
            const uint32_t ae = mapa_rank(smem_u32(&acc_empty[ab]), 0);
            uint32_t v[32];
            (void)v;
            if (MODE == 0) {
                // bound of the row maximum -> power-of-two scale of the fp16 pair; consumers derive the same exponent
                int i1 = 0, i2 = 0;
                float bound = 0.f;
                if (ok) {
                    i1 = __ldg(p.src + mrow);
                    i2 = __ldg(p.dst + mrow);
                    bound = __fadd_rn(__fadd_rn(__ldg(p.wf_bound), __ldg(p.amax_pq + i1)), __ldg(p.amax_pq + i2));
                    if (n0 == 0 && cg == 0) p.bound_out[mrow] = bound;
                }
                const float osc = pow2f(-exp8(bound));
                const float* prow = p.P + (long long)i1 * p.ld_pq + n0 + cg * WCOLS;
                const float* qrow = p.Q + (long long)i2 * p.ld_pq + n0 + cg * WCOLS;
                // the first group's gathered rows are fetched before the accumulators are complete (their L2 latency hides behind
                // the end of the main loop); the other groups' behind the TMEM read (one group ahead for all of them spilled)
                float gp[16], gq[16];
                ldnc8(prow, gp);
                ldnc8(prow + 8, gp + 8);
                ldnc8(qrow, gq);
                ldnc8(qrow + 8, gq + 8);
                mbar_wait(&acc_full[ab], ((uint32_t)tl >> 1) & 1);
                if (threadIdx.x == 64) ETRACE(tl, 6);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
                for (int gi = 0; gi < WCOLS / 16; ++gi) {
                    const int n = n0 + cg * WCOLS + gi * 16;
                    uint32_t v16[16];
                    tmem_ld16(tbase + (uint32_t)(gi * 16), v16);
                    if (gi > 0) {
                        ldnc8(prow + gi * 16, gp);
                        ldnc8(prow + gi * 16 + 8, gp + 8);
                        ldnc8(qrow + gi * 16, gq);
                        ldnc8(qrow + gi * 16 + 8, gq + 8);
                    }
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    if (gi == WCOLS / 16 - 1) {         // all of this warp's TMEM reads are done: release the accumulators
                        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                        __syncwarp();
                        if (lane == 0) mbar_arrive_cluster(ae);
                    }
                    uint32_t hi[8], lo[8];
#pragma unroll
                    for (int j = 0; j < 16; j += 4) {
                        const float4 cs = __ldg(reinterpret_cast<const float4*>(p.col_scale + n + j));
                        const float c4[4] = {cs.x, cs.y, cs.z, cs.w};
                        float a[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u)
                            a[u] = silu_fast(fmaf(p.alpha * __uint_as_float(v16[j + u]), c4[u], gp[j + u] + gq[j + u])) * osc;
                        split2<1>(a[0], a[1], hi[j / 2], lo[j / 2]);
                        split2<1>(a[2], a[3], hi[j / 2 + 1], lo[j / 2 + 1]);
                    }
                    // The group leaves through TMA: a row-per-lane global store costs one L1 wavefront per lane and instruction
                    // (a third of this epilogue's L1 time when it paced the kernel: scripts/trace_edge.py); 16-byte swizzled
                    // shared-memory stores cost 4.
                    uint8_t* sh = reinterpret_cast<uint8_t*>(ebuf_all) + (warp - EPI_WARP0) * 2048;
                    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the previous group has left the buffer
                    __syncwarp();
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        const int off = lane * 32 + ((c ^ ((lane >> 2) & 1)) << 4);      // SWIZZLE_32B: 16-byte chunk ^ (row / 4) % 2
                        *reinterpret_cast<uint4*>(sh + off) = make_uint4(hi[4 * c], hi[4 * c + 1], hi[4 * c + 2], hi[4 * c + 3]);
                        *reinterpret_cast<uint4*>(sh + 1024 + off) = make_uint4(lo[4 * c], lo[4 * c + 1], lo[4 * c + 2], lo[4 * c + 3]);
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) {
                        const int r0_ = m0 + (int)rank * TM + q * 32;
                        tma_store_2d(&mapOh, sh, n, r0_);
                        tma_store_2d(&mapOl, sh + 1024, n, r0_);
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                }
                if (tl == my_tiles - 1 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
            } else {
                const float rowsc = p.alpha * pow2f(exp8(ok ? __ldg(p.a_bound + mrow) : 0.f));
                int my_seg = -1;
                float my_w = 0.f;
                if (ok) {
                    my_seg = __ldg(p.scat_idx + mrow);
                    my_w = __ldg(p.scat_w + mrow);
                }
                // bit r: row r of this warp's 32-row window starts a segment (row 0 always does); rows past M form a last
                // "segment" with index -1 that is skipped
                const int prev = __shfl_up_sync(0xffffffffu, my_seg, 1);
                const uint32_t seg_starts = __ballot_sync(0xffffffffu, lane == 0 || my_seg != prev);
                float seg_rmax = 0.f;
                mbar_wait(&acc_full[ab], ((uint32_t)tl >> 1) & 1);
                if (threadIdx.x == 64) ETRACE(tl, 6);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                tmem_ld32(tbase, v);
#pragma unroll 1
                for (int cc = 0; cc < 4; ++cc) {
                    const int nb = n0 + cg * 128 + cc * 32;
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int j = 0; j < 32; j += 2) {
                        const float2 cs2 = __ldg(reinterpret_cast<const float2*>(p.col_scale + nb + j));
                        const float2 b2 = p.bias ? __ldg(reinterpret_cast<const float2*>(p.bias + nb + j)) : make_float2(0.f, 0.f);
                        const float y0 = silu_fast(fmaf(rowsc * __uint_as_float(v[j]), cs2.x, b2.x));
                        const float y1 = silu_fast(fmaf(rowsc * __uint_as_float(v[j + 1]), cs2.y, b2.y));
                        seg_rmax = fmaxf(seg_rmax, fmaxf(fabsf(y0), fabsf(y1)));
                        *reinterpret_cast<float2*>(ebuf + lane * EP + j) = make_float2(y0, y1);
                    }
                    __syncwarp();
                    if (cc + 1 < 4) tmem_ld32(tbase + (uint32_t)((cc + 1) * 32), v);      // in flight behind the segment sums
                    else {
                        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                        __syncwarp();
                        if (lane == 0) mbar_arrive_cluster(ae);
                    }
                    scatter_chunk(ebuf, seg_starts, my_seg, my_w, p.scat_out, p.scat_ld, nb, lane);
                    __syncwarp();
                }
                if (p.scat_amax && my_seg >= 0) atomicMax(reinterpret_cast<unsigned*>(p.scat_amax + my_seg), __float_as_uint(seg_rmax));
            }
            if (threadIdx.x == 64) ETRACE(tl, 7);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_pair();                                          // both CTAs are done with the pair's TMEM and barriers
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

template <int MODE>
int launch_pair(const void* A_hi, const void* A_lo, int lda, const void* W_hi, const void* W_lo, int ldw, const EdgeParams& p,
                cudaStream_t s) {
    using C = ECfg<MODE>;
    int rc = mi_tc_get_encode();
    if (rc != MI_OK) return rc;
    CUtensorMap mAh, mAl, mWh, mWl;
    if ((rc = mi_tc_make_map(&mAh, A_hi, p.M, p.K, lda, TM, true)) != MI_OK) return rc;
    if ((rc = mi_tc_make_map(&mAl, A_lo, p.M, p.K, lda, TM, true)) != MI_OK) return rc;
    if ((rc = mi_tc_make_map(&mWh, W_hi, p.N, p.K, ldw, TNP / 2, true)) != MI_OK) return rc;
    if ((rc = mi_tc_make_map(&mWl, W_lo, p.N, p.K, ldw, TNP / 2, true)) != MI_OK) return rc;
    CUtensorMap mOh = mAh, mOl = mAl;                             // block 2 has no tensor output
    if (MODE == 0) {
        if ((rc = mi_tc_make_map_h16(&mOh, p.out_hi, p.M, p.N, p.ld_out, 32)) != MI_OK) return rc;
        if ((rc = mi_tc_make_map_h16(&mOl, p.out_lo, p.M, p.N, p.ld_out, 32)) != MI_OK) return rc;
    }
    static bool attr = false;
    if (!attr) {
        MI_CUDA(cudaFuncSetAttribute(edge_pair_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
        attr = true;
    }
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        MI_CUDA(cudaGetDevice(&dev));
        MI_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    }
    const long long tiles = (long long)mi_div_up(p.M, 2 * TM) * (p.N / TNP);
    long long pairs = sms / 2;
    if (tiles < pairs) pairs = tiles;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(2 * pairs));
    cfg.blockDim = dim3(C::THREADS);
    cfg.dynamicSmemBytes = C::SMEM_BYTES;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = mi_pdl_enabled() ? 1 : 0;
    MI_CUDA(cudaLaunchKernelEx(&cfg, edge_pair_kernel<MODE>, mAh, mAl, mWh, mWl, mOh, mOl, p));
    MI_CHECK_LAUNCH();
    return MI_OK;
}

bool al32(const void* q) { return (reinterpret_cast<uintptr_t>(q) & 31) == 0; }

}  // namespace

#ifdef MI_EDGE_TRACE
extern "C" int mi_edge_trace_read(long long* out, int n) {
    cudaDeviceSynchronize();
    return (int)cudaMemcpyFromSymbol(out, g_etrace, sizeof(long long) * (size_t)n);
}
#endif

extern "C" int mi_edge_block1(int E, int N, int K, const void* phi_hi, const void* phi_lo, int ld_phi, const void* w_hi,
                              const void* w_lo, int ld_w, const float* col_scale, float alpha, const float* P, const float* Q,
                              int ld_pq, const int* src, const int* dst, const float* amax_pq, const float* wf_bound, void* a_hi,
                              void* a_lo, int ld_a, float* a_bound, mi_stream_t stream) {
    MI_CHECK_ARG(E >= 0 && N > 0 && K > 0 && N % TNP == 0, "bad dimension (N must be a multiple of 256)");
    if (E == 0) return MI_OK;
    MI_CHECK_ARG(phi_hi && phi_lo && w_hi && w_lo && col_scale && P && Q && src && dst && amax_pq && wf_bound && a_hi && a_lo &&
                 a_bound, "null pointer");
    MI_CHECK_ARG(ld_phi >= K && ld_w >= K && ld_a >= N && ld_pq >= N, "leading dimension too small");
    MI_CHECK_ARG(ld_phi % 8 == 0 && ld_w % 8 == 0 && mi_host_aligned16(phi_hi) && mi_host_aligned16(phi_lo) &&
                 mi_host_aligned16(w_hi) && mi_host_aligned16(w_lo), "TMA operands need 16-byte aligned rows");
    MI_CHECK_ARG(ld_pq % 8 == 0 && ld_a % 16 == 0 && al32(P) && al32(Q) && al32(a_hi) && al32(a_lo) && mi_host_aligned16(col_scale),
                 "the epilogue needs 32-byte aligned rows");
    EdgeParams p = {};
    p.M = E; p.N = N; p.K = K; p.alpha = alpha; p.col_scale = col_scale;
    p.P = P; p.Q = Q; p.ld_pq = ld_pq; p.src = src; p.dst = dst; p.amax_pq = amax_pq; p.wf_bound = wf_bound;
    p.out_hi = (__half*)a_hi; p.out_lo = (__half*)a_lo; p.ld_out = ld_a; p.bound_out = a_bound;
    return launch_pair<0>(phi_hi, phi_lo, ld_phi, w_hi, w_lo, ld_w, p, (cudaStream_t)stream);
}

extern "C" int mi_edge_block2(int E, int N, int K, const void* a_hi, const void* a_lo, int ld_a, const float* a_bound,
                              const void* w_hi, const void* w_lo, int ld_w, const float* col_scale, const float* bias,
                              float* scat_out, int scat_ld, const int* scat_idx, const float* scat_w, float* scat_amax,
                              mi_stream_t stream) {
    MI_CHECK_ARG(E >= 0 && N > 0 && K > 0 && N % TNP == 0, "bad dimension (N must be a multiple of 256)");
    if (E == 0) return MI_OK;
    MI_CHECK_ARG(a_hi && a_lo && a_bound && w_hi && w_lo && col_scale && scat_out && scat_idx && scat_w, "null pointer");
    MI_CHECK_ARG(ld_a >= K && ld_w >= K && scat_ld >= N, "leading dimension too small");
    MI_CHECK_ARG(ld_a % 8 == 0 && ld_w % 8 == 0 && mi_host_aligned16(a_hi) && mi_host_aligned16(a_lo) && mi_host_aligned16(w_hi) &&
                 mi_host_aligned16(w_lo), "TMA operands need 16-byte aligned rows");
    MI_CHECK_ARG(mi_host_aligned16(col_scale) && (!bias || mi_host_aligned16(bias)), "col_scale / bias must be 16-byte aligned");
    EdgeParams p = {};
    p.M = E; p.N = N; p.K = K; p.alpha = 1.0f; p.col_scale = col_scale;
    p.a_bound = a_bound; p.bias = bias;
    p.scat_out = scat_out; p.scat_ld = scat_ld; p.scat_idx = scat_idx; p.scat_w = scat_w; p.scat_amax = scat_amax;
    return launch_pair<1>(a_hi, a_lo, ld_a, w_hi, w_lo, ld_w, p, (cudaStream_t)stream);
}
