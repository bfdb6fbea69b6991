// Tensor-core path for the dense blocks: FP32-grade GEMM by split-precision FP16 on tcgen05.
//
//   C[M,N] = epilogue( A[M,K] . W[N,K]^T ),  A, W fp32 row-major (K contiguous), fp32 accumulate in TMEM.
//
// 1e-4 parity after 2 000 chained score-network evaluations rules out plain TF32/BF16/FP16 inputs, so every
// operand is split into an fp16 head and an fp16 tail (22+ mantissa bits) and the product is formed from three
// tcgen05.mma kind::f16 per k-slice of 16: a_hi.w_hi + a_lo.w_hi + a_hi.w_lo.  Same accuracy as the classic 3xTF32
// scheme at twice the MMA rate and half the operand bytes.  Two operand formats (include/matinvent_b200.h):
//   * two accumulators (main, 2^11-scaled correction), tail scaled by 2^11, 128-wide tiles: the tensor core
//     truncates when it adds into the accumulator, so the small terms stay out of the main accumulator and are
//     folded in by the epilogue in fp32 (lowest error);
//   * "merged": one accumulator, unscaled tail, both operands pre-scaled by powers of two into [2^14, 2^15) — a
//     128x256 tile then double-buffers in TMEM and moves 1.47x fewer shared-memory bytes per flop.
// fp32 dynamic range: rows of A are rescaled by the power of two their producer's row maximum calls for (exact) and
// the result row is scaled back in the epilogue; W rows carry per-row scales undone by epi->col_scale.
//   * W_hi / W_lo are split once per weight update and streamed by TMA (SWIZZLE_64B rows);
//   * A is streamed by TMA as raw fp32 (SWIZZLE_128B rows) into its own ring and split into two fp16 tiles of an
//     operand slot by the split warps, or arrives pre-split from its producer (mi_tc_gemm_presplit);
//   * fused epilogue (column scales, bias, row gathers, pre-activation store, SiLU or SiLU' of the backward, residual,
//     row maxima for the next GEMM), same contract as mi_sgemm; epi->splitk > 1 cuts K into parts that are ADDED to C
//     with 16-byte reductions (weight gradients).
//
// Persistent CTA per SM, 128 x TN output tiles (TN in {256 merged, 128, 64}): w0 TMA producer, w1 MMA issuer + TMEM
// owner, split warps, epilogue warps (overlapped with the next tile's main loop: TMEM holds the accumulators of two
// tiles).  What bounds it and what was tried: profiles/r1_tc_trace.md.
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>
#include <cudaTypedefs.h>

#include "mi_tc_common.cuh"

using namespace mi_tc;

namespace {

constexpr int TM = 128;                              // CTA tile rows; columns TN in {256, 128, 64} (template)
// Warp roles: w0 TMA, w1 MMA, then
//   A split in the kernel : w2..7 operand split, w8..15 epilogue (16 warps: 128 registers per thread)
//   A pre-split (PRESPLIT): w2..3 idle, w4..19 epilogue — sixteen epilogue warps, because that kernel's epilogue carries the
//                           row gathers of the first per-edge block and is latency-bound with eight (96 registers per thread)
constexpr int SPLIT_THREADS = 192;
constexpr int TK = 32;                               // k-block: 32 elements = 128 B of fp32, 64 B of fp16
// MERGED = 0: two accumulators per tile (main, 2^11-scaled correction), TN <= 128.
// MERGED = 1: one accumulator per tile, operands carry an UNSCALED fp16 tail (both operands are pre-scaled by
//             powers of two into [2^14, 2^15) so the tail stays in fp16's useful range): TN = 256 fits the double
//             buffer, halving operand bytes and split work per flop at ~2.7x the (still FP32-grade) rounding error,
//             because three times as many truncating accumulations go into the one accumulator.
// Shared memory holds two rings, so that the HBM latency of A does not sit on the round trip of an MMA operand slot
// and a pre-split A needs no raw staging at all.  The main loop is bound by shared-memory bandwidth, not by the
// tensor core (profiles/r1_tc_trace.md): per 32-wide k-block of a 128x256 tile the UMMAs read 72 KB of operands,
// TMA writes 48 KB and the split warps move 32 KB, 152 KB at 128 B/clk = 1190 cycles against 768 of MMA issue.
//   raw ring  R x 16 KB : fp32 A tiles as TMA delivers them (SWIZZLE_128B); freed as soon as the split warps read them
//   op ring   S x OPB   : A_hi 8K | A_lo 8K | W_hi TN*64 | W_lo TN*64 (fp16, SWIZZLE_64B); freed by tcgen05.commit
// PRESPLIT (A arrives as fp16 hi/lo from its producer): no raw ring, deeper op ring.
template <int TN, int MERGED, int PRESPLIT>
struct Cfg {
    static constexpr int A_RAW = TM * TK * 4;                     // 16 KB fp32 tile
    static constexpr int A_H = TM * TK * 2;                       // 8 KB fp16 tile, x2 (hi, lo)
    static constexpr int W_H = TN * TK * 2;                       // fp16 weight tile, x2
    static constexpr int OPB = 2 * A_H + 2 * W_H;                 // one op-ring slot
    static constexpr int EPI_WARPS = PRESPLIT ? 16 : 8;
    static constexpr int EPI_WARP0 = PRESPLIT ? 4 : 8;            // first epilogue warp
    static constexpr int THREADS = (EPI_WARP0 + EPI_WARPS) * 32;
    static constexpr int EPITCH = 34;                             // floats per transpose-buffer row (float2 accesses, conflict-free)
    static constexpr int EBUF_BYTES = EPI_WARPS * 32 * EPITCH * 4;   // per-warp transpose buffers of the epilogue
    static constexpr int BAR_BYTES = 256;
    static constexpr int RING_BUDGET = 232448 - EBUF_BYTES - BAR_BYTES - 128;
    static constexpr int R = PRESPLIT ? 0 : (TN == 256 ? 3 : 4);
    static constexpr int S_FIT = (RING_BUDGET - R * A_RAW) / OPB;
    static constexpr int S = S_FIT > 6 ? 6 : S_FIT;
    static constexpr int RAW_BYTES = R * A_RAW, OP_BYTES = S * OPB;
    static constexpr int SMEM_BYTES = RAW_BYTES + OP_BYTES + EBUF_BYTES + BAR_BYTES + 128 /*row exponents*/;
    static constexpr uint32_t ACC_COLS = MERGED ? TN : 2 * TN;    // accumulator columns of one tile
    static constexpr uint32_t TMEM_COLS = 2 * ACC_COLS;           // double buffered: tile t+1 accumulates while t drains
    // tcgen05 instruction descriptor, kind::f16: D=f32 (bits 4-5 = 1), A=B=f16 (bits 7-9, 10-12 = 0),
    // both K-major (bits 15,16 = 0), N>>3 at bits 17-22, M>>4 at bits 24-28.
    static constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
    static_assert(S >= 2 && SMEM_BYTES <= 232448 && TMEM_COLS <= 512, "configuration does not fit the SM");
    static_assert((3 * S + 2 * R + 4) * 8 + 8 <= BAR_BYTES, "barrier block too small");
};

#ifdef MI_TC_TRACE
// Developer instrumentation (scripts/trace_tc.py builds a separate library with -DMI_TC_TRACE; never in the product
// build): per-CTA, per-tile SM clock stamps of the pipeline roles.
constexpr int TR_TILES = 8, TR_SLOTS = 24;
__device__ long long g_trace[160 * TR_TILES * TR_SLOTS];
#define TRACE(tileidx, slot)                                                                                   \
    do {                                                                                                       \
        if ((tileidx) < TR_TILES) g_trace[(blockIdx.x * TR_TILES + (tileidx)) * TR_SLOTS + (slot)] = clock64(); \
    } while (0)
#else
#define TRACE(tileidx, slot) do {} while (0)
#endif

struct TcParams {
    int M, N, K;
    float* C; int ldc;
    mi_epilogue_t e;
    int c_vec;
    int presplit; // A is given as two fp16 arrays (hi, scaled lo): TMA loads them straight into the operand tiles
    int ksplit;   // > 1: K is cut into ksplit parts handled by different work units; results are ADDED to C with 16-byte
                  // reductions (weight gradients: few output tiles, tens of thousands of reduction rows)
};

// Segment sums of one 32x32 chunk held in a warp's transpose buffer (row-major, pitch EP): lane = column, the rows of
// every segment are walked in order and one 128-byte reduction per (segment, chunk) goes to the zeroed destination.  A
// segment (<= ~20 rows with fc edges) meets at most two 32-row windows, so every destination element receives at most
// two partial sums: the result does not depend on their order.
__device__ __forceinline__ void scatter_chunk(const float* ebuf, int EP, uint32_t starts, int my_seg, float my_w,
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

// EPI bit 0: row gathers g1 / g2 present, bit 1: pre-activation store (training), bit 2: third gather g3 and / or
// residual present; EPI = 8: SiLU' epilogue of the backward (v * silu'(z_in)), nothing else.  Compile-time so that the epilogue of the hot instantiations keeps every load of a batch in
// flight (with all four operand sets live the register allocator serialised the gathers: 15k cycles per 32x32 chunk).
// bias / SiLU / column scales stay runtime flags.
template <int TN, int EPI, int MERGED, int PRESPLIT>
__global__ void __launch_bounds__((Cfg<TN, MERGED, PRESPLIT>::THREADS), 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapAlo,
               const __grid_constant__ CUtensorMap mapWhi, const __grid_constant__ CUtensorMap mapWlo, const TcParams p) {
    // Persistent CTA: tiles blockIdx.x, blockIdx.x + gridDim.x, ...  (n fastest, so the CTAs that share an A row
    // block run at the same time and hit it in L2).  All pipeline counters run across tiles (one flat k-block index
    // per CTA), so the rings already stream the next tile while this tile's epilogue drains TMEM.
    using C = Cfg<TN, MERGED, PRESPLIT>;
    constexpr uint32_t TMEM_COLS = C::TMEM_COLS, IDESC = C::IDESC;
    constexpr uint32_t CORR = MERGED ? 0 : TN;            // column offset of the correction accumulator
    constexpr int A_RAW = C::A_RAW, A_H = C::A_H, W_H = C::W_H, OPB = C::OPB, S = C::S, R = C::R;
    constexpr int RR = R > 0 ? R : 1;                     // divisor only
    extern __shared__ __align__(1024) uint8_t smem[];     // swizzled tiles need 1024-byte alignment (checked below)
#ifdef MI_TC_TRACE
    if (threadIdx.x == 0) {                                // wall-clock (ns) at kernel entry: launch ramp vs in-kernel time
        unsigned long long gt;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        g_trace[(blockIdx.x * TR_TILES) * TR_SLOTS + 21] = (long long)gt;
    }
#endif
    uint8_t* raw_ring = smem;
    uint8_t* op_ring = smem + C::RAW_BYTES;
    float* ebuf_all = reinterpret_cast<float*>(smem + C::RAW_BYTES + C::OP_BYTES);      // EPI_WARPS x 32 x EPITCH floats
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::RAW_BYTES + C::OP_BYTES + C::EBUF_BYTES);
    uint64_t* w_full = bars;                       // [S] TMA landed in the op slot (W, and A hi/lo when PRESPLIT)
    uint64_t* a_ready = bars + S;                  // [S] split warps have written A hi/lo
    uint64_t* op_empty = bars + 2 * S;             // [S] MMAs done with the slot
    uint64_t* raw_full = bars + 3 * S;             // [R] fp32 A tile landed
    uint64_t* raw_empty = bars + 3 * S + R;        // [R] split warps have consumed it
    uint64_t* acc_full = bars + 3 * S + 2 * R;     // [2] accumulators of a tile complete
    uint64_t* acc_empty = acc_full + 2;            // [2] epilogue has drained the accumulator buffer
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 4);
    int8_t* rexp = reinterpret_cast<int8_t*>(smem + C::RAW_BYTES + C::OP_BYTES + C::EBUF_BYTES + C::BAR_BYTES);   // [128] row exponents

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // k-blocks per work unit; with split-K every part gets the same count and the last one runs past K, where TMA
    // zero-fills (at most one wasted k-block per part)
    const int nkb = ((p.K + TK - 1) / TK + p.ksplit - 1) / p.ksplit;
    const int tiles_n = (p.N + TN - 1) / TN;
    const int tiles_m = (p.M + TM - 1) / TM;
    const int num_tiles = tiles_n * tiles_m * p.ksplit;               // work units: (tile, K part), part fastest
    const int my_tiles = (num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const uint32_t total = (uint32_t)my_tiles * (uint32_t)nkb;        // k-blocks this CTA goes through
    auto tile_origin4 = [&](int tl, int& m0, int& n0, int& kb0) {
        const int unit = (int)blockIdx.x + tl * (int)gridDim.x;
        const int tile = unit / p.ksplit;
        kb0 = (unit % p.ksplit) * nkb;
        m0 = (tile / tiles_n) * TM;
        n0 = (tile % tiles_n) * TN;
    };
    auto tile_origin = [&](int tl, int& m0, int& n0) {
        int kb0;
        tile_origin4(tl, m0, n0, kb0);
    };

    if (threadIdx.x == 0) {
        if (smem_u32(smem) & 1023u) __trap();
        for (int s = 0; s < S; ++s) {
            mbar_init(&w_full[s], 1);
            mbar_init(&a_ready[s], SPLIT_THREADS / 32);     // one arrive per split warp
            mbar_init(&op_empty[s], 1);
        }
        for (int r = 0; r < R; ++r) {
            mbar_init(&raw_full[r], 1);
            mbar_init(&raw_empty[r], SPLIT_THREADS / 32);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&acc_full[b], 1);
            mbar_init(&acc_empty[b], C::EPI_WARPS);         // one arrive per epilogue warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapAlo) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapWhi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapWlo) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        // One thread feeds both rings: W (and pre-split A) for k-block i as soon as its op slot is free, raw A for
        // k-block i + R - 1 as soon as the split warps have released that raw slot — raw tiles run ahead of the
        // op ring, which is what hides the HBM latency of A.
        // (All lanes run the loops, one elected lane issues: inside an `if (lane == 0)` region the compiler wraps every TMA /
        // MMA instruction in a R2UR.BROADCAST loop — see elect_one() in mi_tc_common.cuh.)
        {
            if (lane == 0) TRACE(0, 14);
            auto issue_raw = [&](uint32_t idx) {
                const int tl = (int)(idx / (uint32_t)nkb), kb = (int)(idx % (uint32_t)nkb);
                int m0, n0, kb0;
                tile_origin4(tl, m0, n0, kb0);
                const int r = (int)(idx % (uint32_t)RR);
                mbar_wait(&raw_empty[r], ((idx / (uint32_t)RR) & 1) ^ 1);
                if (elect_one()) {
                    mbar_expect_tx(&raw_full[r], A_RAW);
                    tma_load_2d(raw_ring + r * A_RAW, &mapA, &raw_full[r], (kb0 + kb) * TK, m0);
                }
                __syncwarp();
            };
            uint32_t a_it = 0;
            if (!PRESPLIT)
                for (; (int)a_it < R - 1 && a_it < total; ++a_it) issue_raw(a_it);
            for (uint32_t it = 0; it < total; ++it) {
                const int tl = (int)(it / (uint32_t)nkb), kb = (int)(it % (uint32_t)nkb);
                int m0, n0, kb0;
                tile_origin4(tl, m0, n0, kb0);
                const int kc = (kb0 + kb) * TK;
                const int s = (int)(it % (uint32_t)S);
                mbar_wait(&op_empty[s], ((it / (uint32_t)S) & 1) ^ 1);
                if (kb == 0 && lane == 0) TRACE(tl, 0);
                uint8_t* st = op_ring + s * OPB;
                if (elect_one()) {
                    mbar_expect_tx(&w_full[s], (PRESPLIT ? 2 * A_H : 0) + 2 * W_H);
                    if (PRESPLIT) {
                        tma_load_2d(st, &mapA, &w_full[s], kc, m0);
                        tma_load_2d(st + A_H, &mapAlo, &w_full[s], kc, m0);
                    }
                    tma_load_2d(st + 2 * A_H, &mapWhi, &w_full[s], kc, n0);
                    tma_load_2d(st + 2 * A_H + W_H, &mapWlo, &w_full[s], kc, n0);
                }
                __syncwarp();
                if (!PRESPLIT && a_it < total) issue_raw(a_it++);
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        {
            uint32_t it = 0;
            for (int tl = 0; tl < my_tiles; ++tl) {
                const uint32_t ab = (uint32_t)tl & 1;            // accumulator buffer of this tile
                const uint32_t acc = tmem_base + ab * C::ACC_COLS;
                mbar_wait(&acc_empty[ab], (((uint32_t)tl >> 1) & 1) ^ 1);   // the tile two back has been read out of this buffer
                if (lane == 0) TRACE(tl, 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int s = (int)(it % (uint32_t)S);
                    const uint32_t ph = (it / (uint32_t)S) & 1;
                    mbar_wait(&w_full[s], ph);
                    if (!PRESPLIT) mbar_wait(&a_ready[s], ph);
                    if (kb == 0 && lane == 0) TRACE(tl, 2);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t st = smem_u32(op_ring + s * OPB);
                    const uint64_t d_ahi = umma_desc(st), d_alo = umma_desc(st + A_H);
                    const uint64_t d_whi = umma_desc(st + 2 * A_H), d_wlo = umma_desc(st + 2 * A_H + W_H);
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < TK / 16; ++k) {
                            const uint64_t adv = (uint64_t)((k * 32) >> 4);      // 16 fp16 = 32 bytes along the swizzled row
                            umma_f16(acc, d_ahi + adv, d_whi + adv, IDESC, (kb | k) != 0);
                            umma_f16(acc + CORR, d_alo + adv, d_whi + adv, IDESC, MERGED ? 1u : (uint32_t)((kb | k) != 0));
                            umma_f16(acc + CORR, d_ahi + adv, d_wlo + adv, IDESC, 1u);
                        }
                        umma_commit(&op_empty[s]);
                        if (kb == nkb - 1) umma_commit(&acc_full[ab]);
                    }
                    __syncwarp();
                }
                if (lane == 0) TRACE(tl, 3);
            }
        }
    } else if (warp < C::EPI_WARP0) {
        // ===================== operand split warps (w2..7; idle when A arrives pre-split) =====================
        const int t = threadIdx.x - 64;     // 0..191
        const mi_epilogue_t& e = p.e;
        uint32_t it = 0;
        for (int tl = 0; tl < my_tiles && !PRESPLIT; ++tl) {
            int m0, n0;
            tile_origin(tl, m0, n0);
            // Row rescaling (fp32 dynamic range on the fp16 tensor path): when the producer of A reports the row
            // maxima, every row is multiplied by the power of two that brings its max |a| into [2^14, 2^15) before
            // the split — exact — and the result row by the inverse in the epilogue.
            {
                int e8 = 0;
                const int m = m0 + t;
                if (e.a_amax && t < TM && m < p.M) {
                    const int ex = (int)((__float_as_uint(__ldg(e.a_amax + m)) >> 23) & 0xff) - 127;
                    e8 = max(-100, min(ex - 14, 100));
                }
                asm volatile("bar.sync 1, %0;" ::"n"(SPLIT_THREADS) : "memory");      // everyone is done reading the previous tile's exponents
                if (t < TM) rexp[t] = (int8_t)e8;
                asm volatile("bar.sync 1, %0;" ::"n"(SPLIT_THREADS) : "memory");
            }
            for (int kb = 0; kb < nkb; ++kb, ++it) {
                const int r = (int)(it % (uint32_t)RR), s = (int)(it % (uint32_t)S);
                mbar_wait(&raw_full[r], (it / (uint32_t)RR) & 1);
                const float4* raw = reinterpret_cast<const float4*>(raw_ring + r * A_RAW);
                uint8_t* hi = op_ring + s * OPB;
                uint8_t* lo = hi + A_H;
                // all loads first (the stores below may alias them as far as the compiler knows), then convert + store
                constexpr int NV = (TM * TK / 4 + SPLIT_THREADS - 1) / SPLIT_THREADS;   // 1024 float4 over 192 threads: 6, last one partial
                float4 v[NV];
                float sc[NV];
#pragma unroll
                for (int i = 0; i < NV; ++i) {
                    const int pidx = i * SPLIT_THREADS + t;           // physical float4 slot in the 128B-swizzled fp32 tile
                    if (pidx < TM * TK / 4) {
                        v[i] = raw[pidx];
                        sc[i] = __uint_as_float((uint32_t)(127 - (int)rexp[pidx >> 3]) << 23);      // 2^-e, exact
                    }
                }
                mbar_wait(&op_empty[s], ((it / (uint32_t)S) & 1) ^ 1);   // MMAs of the slot's previous use are done
#pragma unroll
                for (int i = 0; i < NV; ++i) {
                    const int pidx = i * SPLIT_THREADS + t;
                    if (pidx >= TM * TK / 4) break;
                    const int row = pidx >> 3;
                    const int k0 = ((pidx & 7) ^ (row & 7)) << 2;     // logical k of the slot (Swizzle<3,4,3>)
                    uint2 h, l;
                    split2<MERGED>(v[i].x * sc[i], v[i].y * sc[i], h.x, l.x);
                    split2<MERGED>(v[i].z * sc[i], v[i].w * sc[i], h.y, l.y);
                    // fp16 tile: 64-byte rows, 16-byte chunk index XOR (row/2)%4 (Swizzle<2,4,3>)
                    const int off = row * 64 + ((((k0 >> 3) ^ (row >> 1)) & 3) << 4) + ((k0 & 7) << 1);
                    *reinterpret_cast<uint2*>(hi + off) = h;
                    *reinterpret_cast<uint2*>(lo + off) = l;
                }
                // Release the raw slot only now: the stores above consumed every loaded register, so the LDS results
                // have landed.  (An arrive issued right behind the LDS instructions overtook them, and TMA refilled
                // the slot under the loads now and then.)
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> async proxy (UMMA)
                __syncwarp();                                                   // every lane's stores and fence are done
                if (lane == 0) {
                    mbar_arrive(&raw_empty[r]);
                    mbar_arrive(&a_ready[s]);
                }
#ifdef MI_TC_TRACE
                if (t == 0 && kb == 0) TRACE(tl, 13);
#endif
            }
        }
    } else {
        // ===================== epilogue warps (w8..15), overlapped with the next tile's main loop =====================
        const int q = warp & 3;                      // TMEM lane quarter this warp may access
        const int cg = (warp - C::EPI_WARP0) >> 2;   // column group handled by this warp
        constexpr int EP = C::EPITCH;
        float* ebuf = ebuf_all + (warp - C::EPI_WARP0) * (32 * EP);
        const mi_epilogue_t& e = p.e;
        constexpr int NCH = TN / 32, NCG = C::EPI_WARPS / 4;     // 32-column chunks of a tile, column groups of warps
        constexpr int CH = NCH >= NCG ? NCH / NCG : 1;           // chunks per warp (narrow tiles leave column groups >= NCH idle)
        for (int tl = 0; tl < my_tiles; ++tl) {
            int m0, n0;
            tile_origin(tl, m0, n0);
            const uint32_t ab = (uint32_t)tl & 1;
            const uint32_t acc = tmem_base + ab * C::ACC_COLS;
            // ---- epilogue: TMEM -> registers -> per-warp smem transpose -> coalesced global traffic
            const int mrow = m0 + q * 32 + lane;                 // the row this lane owns in TMEM
            int e8 = 0;                                          // same power-of-two row exponent as the split warps
            if (e.a_amax && mrow < p.M) {
                const int ex = (int)((__float_as_uint(__ldg(e.a_amax + mrow)) >> 23) & 0xff) - 127;
                e8 = max(-100, min(ex - 14, 100));
            }
            const float rowsc = e.alpha * __uint_as_float((uint32_t)(127 + e8) << 23);   // alpha * 2^e
            int i1 = 0, i2 = 0, i3 = 0;
            if (EPI & 1) {
                const bool mrow_ok = mrow < p.M;
                i1 = (mrow_ok && e.g1) ? (e.g1_idx ? __ldg(e.g1_idx + mrow) : mrow) : 0;
                i2 = (mrow_ok && e.g2) ? (e.g2_idx ? __ldg(e.g2_idx + mrow) : mrow) : 0;
            }
            if (EPI & 4) {
                const bool mrow_ok = mrow < p.M;
                i3 = (mrow_ok && e.g3) ? (e.g3_idx ? __ldg(e.g3_idx + mrow) : mrow) : 0;
            }
            // fused scatter-mean (EPI bit 4): the segment (destination row of the reduction) and the weight 1 / count of
            // the row this lane owns in TMEM; rows of one segment are consecutive
            int my_seg = -1;
            float my_w = 0.f;
            if ((EPI & 16) && mrow < p.M) {
                my_seg = __ldg(e.scat_idx + mrow);
                my_w = __ldg(e.scat_w + mrow);
            }
            // bit r: row r of this warp's 32-row window starts a segment (row 0 always does); rows past M form a last
            // "segment" with index -1 that is skipped
            uint32_t seg_starts = 0;
            float seg_rmax = 0.f;                                // max |v| of the row this lane owns, all columns (EPI == 16)
            if (EPI & 16) {
                const int prev = __shfl_up_sync(0xffffffffu, my_seg, 1);
                seg_starts = __ballot_sync(0xffffffffu, lane == 0 || my_seg != prev);
            }
            float rowmax[8];                                     // running max |C| of the 8 rows this lane touches
#pragma unroll
            for (int u = 0; u < 8; ++u) rowmax[u] = 0.f;
            mbar_wait(&acc_full[ab], ((uint32_t)tl >> 1) & 1);
            if (threadIdx.x == 320) TRACE(tl, 4);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            // Next chunk's TMEM read in flight while this one is written out — only where the registers allow it
            // (one accumulator, no gather operands).
            constexpr bool PREFETCH = MERGED && !(EPI & 1);
            uint32_t v[32], w[32];
            const uint32_t tbase = acc + ((uint32_t)(q * 32) << 16) + (uint32_t)(cg * CH * 32);
            if (PREFETCH) tmem_ld32(tbase, v);
#pragma unroll 1
            for (int cc = 0; cc < CH; ++cc) {
                const int nb = n0 + (cg * CH + cc) * 32;
                const int col4 = (lane & 7) * 4;
                const int n = nb + col4;
                const bool live = nb < p.N && cg * CH + cc < NCH;      // warp-uniform
                const bool fast = live && p.c_vec && nb + 32 <= p.N;
                // Row gathers g1 / g2 of the chunk: 2 x 4 x 16-byte loads per lane and half chunk, issued so that their
                // L2 round trips hide behind the TMEM read + transpose (first half) and behind the first half's
                // arithmetic (second half).  Rows past M read row index 0 and are dropped at the store.
                float4 ga0[4], gb0[4], ga1[4], gb1[4];
                auto issue_gathers = [&](int hb, float4 (&ga)[4], float4 (&gb)[4]) {
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int rr = (hb * 4 + u) * 4 + (lane >> 3);
                        const int r1 = __shfl_sync(0xffffffffu, i1, rr), r2 = __shfl_sync(0xffffffffu, i2, rr);
                        ga[u] = gb[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (e.g1) ga[u] = ldg_stream4(e.g1 + (long long)r1 * e.g1_ld + n);
                        if (e.g2) gb[u] = ldg_stream4(e.g2 + (long long)r2 * e.g2_ld + n);
                    }
                };
                if (!PREFETCH && cg * CH + cc < NCH) {
                    tmem_ld32(tbase + (uint32_t)(cc * 32), v);
                    if (!MERGED) tmem_ld32(tbase + (uint32_t)(cc * 32) + CORR, w);
                }
                if ((EPI & 1) && fast) issue_gathers(0, ga0, gb0);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (threadIdx.x == 320 && cc < 4) TRACE(tl, 5 + cc);
                if (cc == CH - 1) {                      // all of this warp's TMEM reads are done: release the accumulators
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&acc_empty[ab]);
                }
                if (EPI == 16) {
                    // inference form of the fused scatter: column scale, bias and SiLU are applied here, in the
                    // row-per-lane layout TMEM delivers (the per-column constants are warp-uniform loads), the values go
                    // through the transpose buffer once and are summed by segment: no second stage, no global store of C
                    if (live) {
#pragma unroll
                        for (int j = 0; j < 32; j += 2) {
                            const float2 cs2 = e.col_scale ? __ldg(reinterpret_cast<const float2*>(e.col_scale + nb + j)) : make_float2(1.f, 1.f);
                            const float2 b2 = e.bias ? __ldg(reinterpret_cast<const float2*>(e.bias + nb + j)) : make_float2(0.f, 0.f);
                            float y0 = fmaf(rowsc * __uint_as_float(v[j]), cs2.x, b2.x);
                            float y1 = fmaf(rowsc * __uint_as_float(v[j + 1]), cs2.y, b2.y);
                            if (e.act == MI_ACT_SILU) y0 = silu_fast(y0), y1 = silu_fast(y1);
                            seg_rmax = fmaxf(seg_rmax, fmaxf(fabsf(y0), fabsf(y1)));
                            *reinterpret_cast<float2*>(ebuf + lane * EP + j) = make_float2(y0, y1);
                        }
                    }
                    __syncwarp();
                    if (PREFETCH && cc + 1 < CH) tmem_ld32(tbase + (uint32_t)((cc + 1) * 32), v);
                    if (live) scatter_chunk(ebuf, EP, seg_starts, my_seg, my_w, e.scat_out, e.scat_ld, nb, lane);
                    __syncwarp();
                    continue;
                }
                if (live) {
#pragma unroll
                    for (int j = 0; j < 32; j += 2) {        // STS.64, bank = (2*lane + j) % 32: conflict-free
                        float y0 = __uint_as_float(v[j]), y1 = __uint_as_float(v[j + 1]);
                        if (!MERGED) {
                            y0 = fmaf(__uint_as_float(w[j]), LO_UNSCALE, y0);
                            y1 = fmaf(__uint_as_float(w[j + 1]), LO_UNSCALE, y1);
                        }
                        *reinterpret_cast<float2*>(ebuf + lane * EP + j) = make_float2(rowsc * y0, rowsc * y1);
                    }
                }
                __syncwarp();
                if (PREFETCH && cc + 1 < CH) tmem_ld32(tbase + (uint32_t)((cc + 1) * 32), v);
                if ((EPI & 1) && fast) issue_gathers(1, ga1, gb1);
                if (threadIdx.x == 320 && cc == 0) TRACE(tl, 16);
                if (fast) {
                    // ---- fast path: whole chunk in range, 16-byte accesses, two batches of four row groups
                    float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f), cs4 = make_float4(1.f, 1.f, 1.f, 1.f);
                    if (e.bias) bias4 = __ldg(reinterpret_cast<const float4*>(e.bias + n));
                    if (e.col_scale) cs4 = __ldg(reinterpret_cast<const float4*>(e.col_scale + n));
#pragma unroll
                    for (int hb = 0; hb < 2; ++hb) {
                        float4 gc[4], gr[4], gz[4];
                        float2 xa[4], xb[4];
                        int mm[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int rr = (hb * 4 + u) * 4 + (lane >> 3);
                            const int m = m0 + q * 32 + rr;
                            mm[u] = m;
                            const bool ok = m < p.M;
                            gc[u] = gr[u] = gz[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                            if ((EPI & 8) && ok) gz[u] = ldg_stream4(e.z_in + (long long)m * e.zin_ld + n);
                            if (EPI & 4) {
                                const int r3 = __shfl_sync(0xffffffffu, i3, rr);
                                if (ok && e.g3) gc[u] = ldg_stream4(e.g3 + (long long)r3 * e.g3_ld + n);
                                if (ok && e.resid) gr[u] = ldg_stream4(e.resid + (long long)m * e.resid_ld + n);
                            }
                            xa[u] = *reinterpret_cast<const float2*>(ebuf + rr * EP + col4);
                            xb[u] = *reinterpret_cast<const float2*>(ebuf + rr * EP + col4 + 2);
                        }
                        if (threadIdx.x == 320 && cc == 0) TRACE(tl, 17 + 2 * hb);
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int m = mm[u];
                            if (m >= p.M) continue;
                            float x[4] = {fmaf(xa[u].x, cs4.x, bias4.x), fmaf(xa[u].y, cs4.y, bias4.y),
                                          fmaf(xb[u].x, cs4.z, bias4.z), fmaf(xb[u].y, cs4.w, bias4.w)};
                            if (EPI & 1) {
                                const float4 a4 = hb ? ga1[u] : ga0[u], b4 = hb ? gb1[u] : gb0[u];
                                x[0] += a4.x + b4.x; x[1] += a4.y + b4.y; x[2] += a4.z + b4.z; x[3] += a4.w + b4.w;
                            }
                            if (EPI & 4) { x[0] += gc[u].x; x[1] += gc[u].y; x[2] += gc[u].z; x[3] += gc[u].w; }
                            if (EPI & 2) *reinterpret_cast<float4*>(e.z_out + (long long)m * e.z_ld + n) = make_float4(x[0], x[1], x[2], x[3]);
                            if (EPI & 8) {                 // backward of SiLU: v * silu'(z_in)
                                x[0] *= dsilu_fast(gz[u].x); x[1] *= dsilu_fast(gz[u].y);
                                x[2] *= dsilu_fast(gz[u].z); x[3] *= dsilu_fast(gz[u].w);
                            } else if (e.act == MI_ACT_SILU) {
#pragma unroll
                                for (int v4 = 0; v4 < 4; ++v4) x[v4] = silu_fast(x[v4]);
                            }
                            if (EPI & 4) { x[0] += gr[u].x; x[1] += gr[u].y; x[2] += gr[u].z; x[3] += gr[u].w; }
#ifndef MI_TC_NOSTORE
                            if (EPI & 16) {          // back into the transpose buffer (this lane's own slots): summed below
                                const int rr = (hb * 4 + u) * 4 + (lane >> 3);
                                *reinterpret_cast<float2*>(ebuf + rr * EP + col4) = make_float2(x[0], x[1]);
                                *reinterpret_cast<float2*>(ebuf + rr * EP + col4 + 2) = make_float2(x[2], x[3]);
                            } else if (p.ksplit > 1)
                                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p.C + (long long)m * p.ldc + n), "f"(x[0]),
                                             "f"(x[1]), "f"(x[2]), "f"(x[3])
                                             : "memory");
                            else
                                *reinterpret_cast<float4*>(p.C + (long long)m * p.ldc + n) = make_float4(x[0], x[1], x[2], x[3]);
#else
                            if (x[0] == 1.2345f) p.C[0] = x[1] + x[2] + x[3];
#endif
                            rowmax[hb * 4 + u] = fmaxf(rowmax[hb * 4 + u], fmaxf(fmaxf(fabsf(x[0]), fabsf(x[1])), fmaxf(fabsf(x[2]), fabsf(x[3]))));
                        }
                        if (threadIdx.x == 320 && cc == 0) TRACE(tl, 18 + 2 * hb);
                    }
                    if (EPI & 16) {
                        __syncwarp();
                        scatter_chunk(ebuf, EP, seg_starts, my_seg, my_w, e.scat_out, e.scat_ld, nb, lane);
                    }
                } else if (live) {
                    // ---- generic path (ragged N or unaligned rows): scalar, bounds-checked
#pragma unroll 1
                    for (int rr0 = 0; rr0 < 32; rr0 += 4) {
                        const int rr = rr0 + (lane >> 3);
                        const int m = m0 + q * 32 + rr;
                        int r1 = 0, r2 = 0, r3 = 0;
                        if (EPI & 1) {
                            r1 = __shfl_sync(0xffffffffu, i1, rr);
                            r2 = __shfl_sync(0xffffffffu, i2, rr);
                        }
                        if (EPI & 4) r3 = __shfl_sync(0xffffffffu, i3, rr);
                        float rmax = 0.f;
                        if (m < p.M) {
                            float* crow = p.C + (long long)m * p.ldc + n;
                            for (int u = 0; u < 4; ++u) {
                                if (n + u >= p.N) continue;
                                float y = ebuf[rr * EP + col4 + u];
                                if (e.col_scale) y *= __ldg(e.col_scale + n + u);
                                if (e.bias) y += __ldg(e.bias + n + u);
                                if (EPI & 1) {
                                    if (e.g1) y += __ldg(e.g1 + (long long)r1 * e.g1_ld + n + u);
                                    if (e.g2) y += __ldg(e.g2 + (long long)r2 * e.g2_ld + n + u);
                                }
                                if ((EPI & 4) && e.g3) y += __ldg(e.g3 + (long long)r3 * e.g3_ld + n + u);
                                if (EPI & 2) e.z_out[(long long)m * e.z_ld + n + u] = y;
                                if (EPI & 8) y *= dsilu_fast(__ldg(e.z_in + (long long)m * e.zin_ld + n + u));
                                else if (e.act == MI_ACT_SILU) y = silu_fast(y);
                                if ((EPI & 4) && e.resid) y += __ldg(e.resid + (long long)m * e.resid_ld + n + u);
                                if (p.ksplit > 1) atomicAdd(crow + u, y); else crow[u] = y;
                                rmax = fmaxf(rmax, fabsf(y));
                            }
                        }
#pragma unroll
                        for (int u = 0; u < 8; ++u)              // static indices keep rowmax[] in registers
                            if (u == (rr0 >> 2)) rowmax[u] = fmaxf(rowmax[u], rmax);
                    }
                }
                __syncwarp();
                if (threadIdx.x == 320 && cc < 4) TRACE(tl, 9 + cc);
            }
            if ((EPI & 16) && e.scat_amax) {
                // max |v| over the rows of a segment and all columns: an upper bound of the row maximum of the means
                if (EPI == 16) {
                    if (my_seg >= 0) atomicMax(reinterpret_cast<unsigned*>(e.scat_amax + my_seg), __float_as_uint(seg_rmax));
                } else {
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        float rmax = rowmax[u];
                        rmax = fmaxf(rmax, __shfl_xor_sync(0xffffffffu, rmax, 1));
                        rmax = fmaxf(rmax, __shfl_xor_sync(0xffffffffu, rmax, 2));
                        rmax = fmaxf(rmax, __shfl_xor_sync(0xffffffffu, rmax, 4));
                        const int sg = __shfl_sync(0xffffffffu, my_seg, u * 4 + (lane >> 3));
                        if ((lane & 7) == 0 && sg >= 0) atomicMax(reinterpret_cast<unsigned*>(e.scat_amax + sg), __float_as_uint(rmax));
                    }
                }
            }
            if (e.amax_out) {                                    // one atomic per row per warp: max over the 8 lanes sharing a row
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    float rmax = rowmax[u];
                    rmax = fmaxf(rmax, __shfl_xor_sync(0xffffffffu, rmax, 1));
                    rmax = fmaxf(rmax, __shfl_xor_sync(0xffffffffu, rmax, 2));
                    rmax = fmaxf(rmax, __shfl_xor_sync(0xffffffffu, rmax, 4));
                    const int m = m0 + q * 32 + u * 4 + (lane >> 3);
                    if ((lane & 7) == 0 && m < p.M) atomicMax(reinterpret_cast<unsigned*>(e.amax_out + m), __float_as_uint(rmax));
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) TRACE(0, 15);
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
#ifdef MI_TC_TRACE
    if (threadIdx.x == 32) {
        unsigned long long gt;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        g_trace[(blockIdx.x * TR_TILES) * TR_SLOTS + 22] = (long long)gt;
    }
#endif
}

__global__ void f16_split_kernel(const float* __restrict__ w, __half* __restrict__ hi, __half* __restrict__ lo, long long n,
                                 float scale, float lo_scale) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float x = w[i] * scale;
    __half h = __float2half_rn(x);
    hi[i] = h;
    lo[i] = __float2half_rn((x - __half2float(h)) * lo_scale);
}

// one warp per row: power-of-two scale from the row maximum, merged-format split (hi = fp16(s w), lo = fp16(s w - hi))
__global__ void f16_split_rows_kernel(const float* __restrict__ w, int rows, int cols, int ld, __half* __restrict__ hi,
                                      __half* __restrict__ lo, float* __restrict__ inv_scale) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* wr = w + (long long)row * ld;
    float mx = 0.f;
    for (int c = lane; c < cols; c += 32) mx = fmaxf(mx, fabsf(wr[c]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    int ex = (int)((__float_as_uint(mx) >> 23) & 0xff) - 127;         // max in [2^ex, 2^(ex+1))
    if (mx == 0.f || ex > 100) ex = 14;                               // all-zero (or non-finite) row: scale 1
    ex = max(ex, -100);
    const float s = __uint_as_float((uint32_t)(127 + 14 - ex) << 23); // 2^(14 - ex), exact
    for (int c = lane; c < cols; c += 32) {
        const float x = wr[c] * s;
        const __half h = __float2half_rn(x);
        hi[(long long)row * ld + c] = h;
        lo[(long long)row * ld + c] = __float2half_rn(x - __half2float(h));
    }
    if (lane == 0) inv_scale[row] = __uint_as_float((uint32_t)(127 - 14 + ex) << 23);
}

PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;

int get_encode() {
    if (g_encode) return MI_OK;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t err = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (err != cudaSuccess || qres != cudaDriverEntryPointSuccess || fn == nullptr) {
        mi_set_error_("cuTensorMapEncodeTiled is unavailable (%s)", cudaGetErrorString(err));
        return MI_ERR_CUDA;
    }
    g_encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
    return MI_OK;
}

// 2-D row-major [rows, cols] (ld elements between rows), box = [box_rows, 32 cols], zero OOB fill;
// fp32 operands use 128-byte swizzle rows, fp16 operands 64-byte rows.
int make_map(CUtensorMap* map, const void* base, long long rows, long long cols, long long ld, int box_rows, bool half) {
    cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t gstr[1] = {(cuuint64_t)ld * (half ? 2 : 4)};
    cuuint32_t box[2] = {(cuuint32_t)TK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode(map, half ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                          const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          half ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        mi_set_error_("cuTensorMapEncodeTiled failed (%d) for [%lld,%lld] ld %lld", (int)r, rows, cols, ld);
        return MI_ERR_CUDA;
    }
    return MI_OK;
}


template <int TN, int EPI, int MERGED, int PRESPLIT>
int launch_tc(int M, int N, int K, const void* A, const void* A_lo, int lda, const void* W_hi, const void* W_lo, int ldw,
              cudaStream_t s, const TcParams& p) {
    using C = Cfg<TN, MERGED, PRESPLIT>;
    static bool attr = false;
    int rc;
    CUtensorMap mA, mAl, mWh, mWl;
    if ((rc = make_map(&mA, A, M, K, lda, TM, PRESPLIT != 0)) != MI_OK) return rc;
    if ((rc = make_map(&mAl, PRESPLIT ? A_lo : A, M, K, lda, TM, PRESPLIT != 0)) != MI_OK) return rc;
    if ((rc = make_map(&mWh, W_hi, N, K, ldw, TN, true)) != MI_OK) return rc;
    if ((rc = make_map(&mWl, W_lo, N, K, ldw, TN, true)) != MI_OK) return rc;
    if (!attr) {
        MI_CUDA(cudaFuncSetAttribute(tc_gemm_kernel<TN, EPI, MERGED, PRESPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
        attr = true;
    }
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        MI_CUDA(cudaGetDevice(&dev));
        MI_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    }
    const long long tiles = (long long)mi_div_up(N, TN) * mi_div_up(M, TM) * p.ksplit;
    const int grid = (int)(tiles < sms ? tiles : sms);        // persistent: one CTA per SM
    tc_gemm_kernel<TN, EPI, MERGED, PRESPLIT><<<grid, C::THREADS, C::SMEM_BYTES, s>>>(mA, mAl, mWh, mWl, p);
    MI_CHECK_LAUNCH();
    return MI_OK;
}

template <int TN, int MERGED>
int dispatch_tc(int epi_mode, bool presplit, int M, int N, int K, const void* A, const void* A_lo, int lda, const void* W_hi,
                const void* W_lo, int ldw, cudaStream_t s, const TcParams& p) {
#define MI_TC_CASE(E)                                                                                              \
    case E:                                                                                                        \
        return presplit ? launch_tc<TN, E, MERGED, 1>(M, N, K, A, A_lo, lda, W_hi, W_lo, ldw, s, p)                \
                        : launch_tc<TN, E, MERGED, 0>(M, N, K, A, A_lo, lda, W_hi, W_lo, ldw, s, p);
    switch (epi_mode) {
        MI_TC_CASE(0)
        MI_TC_CASE(1)
        MI_TC_CASE(2)
        MI_TC_CASE(3)
        MI_TC_CASE(4)
        MI_TC_CASE(5)
        MI_TC_CASE(6)
        MI_TC_CASE(7)
        default:
        MI_TC_CASE(8)
    }
#undef MI_TC_CASE
}

// fused scatter-mean epilogue (EPI bit 4): the second per-edge block only — merged 128x256 tiles, A split in the kernel
int dispatch_tc_scatter(int epi_mode, int M, int N, int K, const void* A, int lda, const void* W_hi, const void* W_lo, int ldw,
                        cudaStream_t s, const TcParams& p) {
    if (epi_mode == 16) return launch_tc<256, 16, 1, 0>(M, N, K, A, nullptr, lda, W_hi, W_lo, ldw, s, p);
    return launch_tc<256, 18, 1, 0>(M, N, K, A, nullptr, lda, W_hi, W_lo, ldw, s, p);
}

// (hi, lo) operand pair as ONE 3-D map [2 planes, rows, cols]: a single TMA operation lands the hi tile followed by the lo
// tile (box = [2, box_rows, 32 columns]); lo must follow hi in memory by a multiple of 16 bytes
int make_map_pair(CUtensorMap* map, const void* hi, const void* lo, long long rows, long long cols, long long ld, int box_rows) {
    const long long plane = (const char*)lo - (const char*)hi;
    if (plane <= 0 || (plane & 15)) {
        mi_set_error_("operand pair: lo must follow hi by a positive multiple of 16 bytes (got %lld)", plane);
        return MI_ERR_ARG;
    }
    cuuint64_t gdim[3] = {(cuuint64_t)cols, (cuuint64_t)rows, 2};
    cuuint64_t gstr[2] = {(cuuint64_t)ld * 2, (cuuint64_t)plane};
    cuuint32_t box[3] = {(cuuint32_t)TK, (cuuint32_t)box_rows, 2};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(hi), gdim, gstr, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        mi_set_error_("cuTensorMapEncodeTiled failed (%d) for the pair [2,%lld,%lld] ld %lld plane %lld", (int)r, rows, cols, ld, plane);
        return MI_ERR_CUDA;
    }
    return MI_OK;
}

}  // namespace

int mi_tc_get_encode() { return get_encode(); }
// 2-D fp16 map with a [box_rows, 16 columns] box (32-byte rows, SWIZZLE_32B): the store tiles of 16-column epilogue groups
int mi_tc_make_map_h16(CUtensorMap* map, const void* base, long long rows, long long cols, long long ld, int box_rows) {
    cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t gstr[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {16u, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        mi_set_error_("cuTensorMapEncodeTiled failed (%d) for the fp16 store map [%lld,%lld] ld %lld", (int)r, rows, cols, ld);
        return MI_ERR_CUDA;
    }
    return MI_OK;
}
int mi_tc_make_map_pair(CUtensorMap* map, const void* hi, const void* lo, long long rows, long long cols, long long ld, int box_rows) {
    return make_map_pair(map, hi, lo, rows, cols, ld, box_rows);
}
int mi_tc_make_map(CUtensorMap* map, const void* base, long long rows, long long cols, long long ld, int box_rows, bool half) {
    return make_map(map, base, rows, cols, ld, box_rows, half);
}

#ifdef MI_TC_TRACE
extern "C" int mi_tc_trace_read(long long* out, int n) {
    cudaDeviceSynchronize();
    return (int)cudaMemcpyFromSymbol(out, g_trace, sizeof(long long) * (size_t)n);
}
#endif

extern "C" int mi_f16_split(const float* w, void* hi, void* lo, long long n, float scale, float lo_scale, mi_stream_t stream) {
    if (n <= 0) return MI_OK;
    MI_CHECK_ARG(w && hi && lo, "null pointer");
    f16_split_kernel<<<mi_div_up(n, 256), 256, 0, (cudaStream_t)stream>>>(w, (__half*)hi, (__half*)lo, n, scale, lo_scale);
    MI_CHECK_LAUNCH();
    return MI_OK;
}

extern "C" int mi_f16_split_rows(const float* w, int rows, int cols, int ld, void* hi, void* lo, float* inv_scale,
                                 mi_stream_t stream) {
    if (rows <= 0 || cols <= 0) return MI_OK;
    MI_CHECK_ARG(w && hi && lo && inv_scale, "null pointer");
    MI_CHECK_ARG(ld >= cols, "leading dimension too small");
    f16_split_rows_kernel<<<mi_div_up(rows, 8), 256, 0, (cudaStream_t)stream>>>(w, rows, cols, ld, (__half*)hi, (__half*)lo, inv_scale);
    MI_CHECK_LAUNCH();
    return MI_OK;
}

static int tc_gemm_impl(int M, int N, int K, const void* A, const void* A_lo, int lda, const void* W_hi, const void* W_lo,
                        int ldw, float* C, int ldc, const mi_epilogue_t* epi, int flags, mi_stream_t stream) {
    const bool merged = (flags & MI_TC_MERGED) != 0;
    MI_CHECK_ARG(M >= 0 && N >= 0 && K > 0, "bad dimension");
    if (M == 0 || N == 0) return MI_OK;
    const bool scat = epi && epi->scat_out != nullptr;
    MI_CHECK_ARG(A && W_hi && W_lo && (C || scat), "null operand");
    MI_CHECK_ARG(lda >= K && ldw >= K && ldc >= N, "leading dimension too small");
    MI_CHECK_ARG(lda % (A_lo ? 8 : 4) == 0 && ldw % 8 == 0 && mi_host_aligned16(A) && (!A_lo || mi_host_aligned16(A_lo)) &&
                 mi_host_aligned16(W_hi) && mi_host_aligned16(W_lo), "TMA operands need 16-byte aligned rows (fp32: ld % 4, fp16: ld % 8)");
    int rc = get_encode();
    if (rc != MI_OK) return rc;
    TcParams p;
    p.M = M; p.N = N; p.K = K; p.C = C; p.ldc = ldc;
    if (epi) p.e = *epi;
    else {
        mi_epilogue_t z = {};
        z.alpha = 1.f; z.splitk = 1;
        p.e = z;
    }
    p.ksplit = p.e.splitk > 1 ? p.e.splitk : 1;
    if (p.ksplit > 1)
        MI_CHECK_ARG(!p.e.bias && !p.e.g1 && !p.e.g2 && !p.e.g3 && !p.e.z_out && !p.e.resid && p.e.act == MI_ACT_NONE &&
                     !p.e.amax_out, "split-K adds alpha * A W^T to C: it takes a plain epilogue (row / column scales only)");
    MI_CHECK_ARG(p.e.beta == 0.f, "the tensor-core path has no accumulate epilogue (pass C as resid, or use mi_sgemm)");
    if (p.e.act == MI_ACT_DSILU)
        MI_CHECK_ARG(p.e.z_in && !p.e.g1 && !p.e.g2 && !p.e.g3 && !p.e.z_out && !p.e.resid,
                     "MI_ACT_DSILU on the tensor-core path needs z_in and a plain epilogue otherwise");
    bool cv = scat || ((ldc % 4 == 0) && mi_host_aligned16(C));
    const mi_epilogue_t& e = p.e;
    if (e.bias) cv = cv && mi_host_aligned16(e.bias);
    if (e.g1) cv = cv && (e.g1_ld % 4 == 0) && mi_host_aligned16(e.g1);
    if (e.g2) cv = cv && (e.g2_ld % 4 == 0) && mi_host_aligned16(e.g2);
    if (e.g3) cv = cv && (e.g3_ld % 4 == 0) && mi_host_aligned16(e.g3);
    if (e.z_out) cv = cv && (e.z_ld % 4 == 0) && mi_host_aligned16(e.z_out);
    if (e.z_in) cv = cv && (e.zin_ld % 4 == 0) && mi_host_aligned16(e.z_in);
    if (e.resid) cv = cv && (e.resid_ld % 4 == 0) && mi_host_aligned16(e.resid);
    if (e.col_scale) cv = cv && mi_host_aligned16(e.col_scale);
    p.c_vec = cv;
    p.presplit = A_lo != nullptr;
    // pre-split A with a_amax: the producer already scaled the rows by the power of two a_amax calls for
    // (mi_layernorm_fwd_split); the epilogue scales the result rows back
    if (merged) MI_CHECK_ARG(p.presplit || p.e.a_amax != nullptr, "the merged format needs the row maxima of A (epi->a_amax)");
    // Column-tile width: 128 (two double-buffered {main, correction} accumulator pairs fill the 512 TMEM columns);
    // 64 only for narrow outputs.
    int tn = (N <= 64) ? 64 : 128;
    const char* force = getenv("MI_TC_TN");
    if (force) tn = atoi(force) <= 64 ? 64 : 128;
    cudaStream_t s = (cudaStream_t)stream;
    const int epi_mode = p.e.act == MI_ACT_DSILU ? 8 : (((p.e.g1 || p.e.g2) ? 1 : 0) | (p.e.z_out ? 2 : 0) | ((p.e.g3 || p.e.resid) ? 4 : 0));
    if (scat) {
        // the rows of C are not stored: segment means go to scat_out (zeroed by the caller), see mi_epilogue_t
        MI_CHECK_ARG(merged && !p.presplit && (epi_mode & ~2) == 0 && p.ksplit == 1 && !p.e.amax_out,
                     "the fused scatter epilogue exists for the merged, in-kernel-split format with a plain / z_out epilogue");
        MI_CHECK_ARG(p.e.scat_idx && p.e.scat_w && N % 32 == 0 && p.c_vec && p.e.scat_ld >= N,
                     "fused scatter: scat_idx / scat_w required, N % 32 == 0, 16-byte aligned epilogue operands");
        return dispatch_tc_scatter(16 | epi_mode, M, N, K, A, lda, W_hi, W_lo, ldw, (cudaStream_t)stream, p);
    }
    if (merged) return dispatch_tc<256, 1>(epi_mode, p.presplit != 0, M, N, K, A, A_lo, lda, W_hi, W_lo, ldw, s, p);
    if (tn == 128) return dispatch_tc<128, 0>(epi_mode, p.presplit != 0, M, N, K, A, A_lo, lda, W_hi, W_lo, ldw, s, p);
    return dispatch_tc<64, 0>(epi_mode, p.presplit != 0, M, N, K, A, A_lo, lda, W_hi, W_lo, ldw, s, p);
}

extern "C" int mi_tc_gemm(int M, int N, int K, const float* A, int lda, const void* W_hi, const void* W_lo, int ldw,
                          float* C, int ldc, const mi_epilogue_t* epi, int flags, mi_stream_t stream) {
    return tc_gemm_impl(M, N, K, A, nullptr, lda, W_hi, W_lo, ldw, C, ldc, epi, flags, stream);
}

extern "C" int mi_tc_gemm_presplit(int M, int N, int K, const void* A_hi, const void* A_lo, int lda, const void* W_hi,
                                   const void* W_lo, int ldw, float* C, int ldc, const mi_epilogue_t* epi, int flags,
                                   mi_stream_t stream) {
    MI_CHECK_ARG(A_lo != nullptr, "null operand");
    return tc_gemm_impl(M, N, K, A_hi, A_lo, lda, W_hi, W_lo, ldw, C, ldc, epi, flags, stream);
}
