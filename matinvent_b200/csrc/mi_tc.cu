// Tensor-core path for the dense blocks: FP32-grade GEMM by split-precision FP16 on tcgen05.
//
//   C[M,N] = epilogue( A[M,K] . W[N,K]^T ),  A, W fp32 row-major (K contiguous), fp32 accumulate in TMEM.
//
// 1e-4 parity after 2 000 chained score-network evaluations rules out plain TF32/BF16/FP16 inputs, so every
// operand is split  x = x_hi + 2^-11 x_lo  with x_hi = fp16(x), x_lo = fp16((x - x_hi) * 2^11)  (22+ mantissa
// bits, residual <= 2^-24 |x|) and the product is formed as  a.w ~= a_hi.w_hi + 2^-11 (a_lo.w_hi + a_hi.w_lo):
// 3 tcgen05.mma kind::f16 per k-slice of 16.  Same accuracy as the classic 3xTF32 scheme at twice the MMA rate
// and half the operand bytes (an earlier 3xTF32 version of this kernel measured 2140 cycles per 32-wide k-block,
// bound by shared-memory bandwidth: git history).  Range: |x| < 65504 (activations and weights of this network
// are O(1..10); out-of-range inputs must use mi_sgemm).
//   * W_hi / W_lo are split once per weight update (mi_f16_split) and streamed by TMA (SWIZZLE_64B rows);
//   * A is streamed by TMA as raw fp32 (SWIZZLE_128B rows) and split into two fp16 tiles in shared memory by the
//     split warpgroup while earlier stages' MMAs run;
//   * two accumulators in TMEM: main (a_hi.w_hi) and correction (scaled 2^11): the tensor core truncates when it
//     adds into the accumulator, so the error grows with the number of accumulating instructions — the small
//     terms stay out of the main accumulator and are folded in by the epilogue in fp32;
//   * fused epilogue (bias + up to 3 row gathers + pre-activation store + SiLU + residual), same contract as
//     mi_sgemm, 8 warps reading TMEM with tcgen05.ld.
//
// Persistent CTA per SM, 128 x TN output tiles (TN in {128,64}), 14 warps: w0 TMA producer, w1 MMA issuer + TMEM
// owner, w2..9 operand split, w10..17 epilogue (overlapped with the next tile's main loop: TMEM holds two
// {main, correction} accumulator pairs).  Stage = A_raw 16K | A_hi 8K | A_lo 8K | W_hi TN*64 | W_lo TN*64, 4 stages.
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>
#include <cudaTypedefs.h>

#include "mi_common.cuh"

namespace {

constexpr int TM = 128;                              // CTA tile rows; columns TN in {128, 64} (template)
constexpr int TC_THREADS = 576;   // w0 TMA, w1 MMA, w2..9 operand split, w10..17 epilogue
constexpr int TK = 32;                               // k-block: 32 elements = 128 B of fp32, 64 B of fp16
// MERGED = 0: two accumulators per tile (main, 2^11-scaled correction), TN <= 128.
// MERGED = 1: one accumulator per tile, operands carry an UNSCALED fp16 tail (both operands are pre-scaled by
//             powers of two into [2^14, 2^15) so the tail stays in fp16's useful range): TN = 256 fits the double
//             buffer, halving operand bytes and split work per flop at ~2.7x the (still FP32-grade) rounding error,
//             because three times as many truncating accumulations go into the one accumulator.
template <int STAGES, int TN, int MERGED = 0>
struct Cfg {
    static constexpr int A_RAW = TM * TK * 4;                     // 16 KB fp32 tile (TMA, SWIZZLE_128B)
    static constexpr int A_H = TM * TK * 2;                       // 8 KB fp16 tile (SWIZZLE_64B), x2 (hi, lo)
    static constexpr int W_H = TN * TK * 2;                       // fp16 weight tile (TMA, SWIZZLE_64B), x2
    static constexpr int STAGE_BYTES = A_RAW + 2 * A_H + 2 * W_H;
    static constexpr int EPITCH = 34;                             // floats per transpose-buffer row (float2 accesses, conflict-free)
    static constexpr int EBUF_BYTES = 8 * 32 * EPITCH * 4;        // per-warp transpose buffers of the epilogue
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EBUF_BYTES + 128 /*barriers*/ + 128 /*row exponents*/;
    static constexpr uint32_t ACC_COLS = MERGED ? TN : 2 * TN;    // accumulator columns of one tile
    static constexpr uint32_t TMEM_COLS = 2 * ACC_COLS;           // double buffered: tile t+1 accumulates while t drains
    // tcgen05 instruction descriptor, kind::f16: D=f32 (bits 4-5 = 1), A=B=f16 (bits 7-9, 10-12 = 0),
    // both K-major (bits 15,16 = 0), N>>3 at bits 17-22, M>>4 at bits 24-28.
    static constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
};
constexpr float LO_SCALE = 2048.0f, LO_UNSCALE = 1.0f / 2048.0f;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
// K-major fp16 operand tile, 64-byte rows, SWIZZLE_64B: 8-row groups of 512 B (SBO), LBO unused (=1),
// descriptor version 1, layout type 4.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(512 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)4 << 61;
    return d;
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// SiLU with ex2/rcp approximations (~3e-7 relative, below the GEMM's own error): 5 instructions instead of ~25,
// the epilogue is issue/latency bound otherwise.
__device__ __forceinline__ float silu_fast(float x) { return __fdividef(x, 1.0f + __expf(-x)); }
// x -> (fp16(x), fp16((x - fp16(x)) * 2^11)) packed for two consecutive elements
template <int MERGED>
__device__ __forceinline__ void split2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
    __half2 h = __floats2half2_rn(x0, x1);
    float2 hf = __half22float2(h);
    const float ls = MERGED ? 1.0f : LO_SCALE;
    __half2 l = __floats2half2_rn((x0 - hf.x) * ls, (x1 - hf.y) * ls);
    hi = *reinterpret_cast<uint32_t*>(&h);
    lo = *reinterpret_cast<uint32_t*>(&l);
}

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
};

// EPI bit 0: row gathers present, bit 1: pre-activation store (training).  bias / SiLU / residual stay runtime flags.
template <int STAGES, int TN, int EPI, int MERGED>
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapAlo,
               const __grid_constant__ CUtensorMap mapWhi, const __grid_constant__ CUtensorMap mapWlo, const TcParams p) {
    // Persistent CTA: tiles blockIdx.x, blockIdx.x + gridDim.x, ...  (n fastest, so the CTAs that share an A row
    // block run at the same time and hit it in L2).  Pipeline counters run across tiles, so the producer already
    // streams the next tile's first stages while this tile's epilogue drains TMEM.
    using C = Cfg<STAGES, TN, MERGED>;
    constexpr uint32_t TMEM_COLS = C::TMEM_COLS, IDESC = C::IDESC;
    constexpr uint32_t CORR = MERGED ? 0 : TN;            // column offset of the correction accumulator
    constexpr int A_RAW = C::A_RAW, A_H = C::A_H, W_H = C::W_H, STAGE_BYTES = C::STAGE_BYTES;
    extern __shared__ __align__(1024) uint8_t smem[];     // swizzled tiles need 1024-byte alignment (checked below)
    float* ebuf_all = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES);          // 8 warps x 32 x 36 floats
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES + C::EBUF_BYTES);
    uint64_t* full = bars;                 // [STAGES] TMA landed
    uint64_t* split = bars + STAGES;       // [STAGES] fp16 A tiles ready
    uint64_t* empty = bars + 2 * STAGES;   // [STAGES] MMAs done with the stage
    uint64_t* acc_full = bars + 3 * STAGES;        // [2] accumulators of a tile complete
    uint64_t* acc_empty = bars + 3 * STAGES + 2;   // [2] epilogue has drained the accumulator buffer
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * STAGES + 4);
    int8_t* rexp = reinterpret_cast<int8_t*>(smem + STAGES * STAGE_BYTES + C::EBUF_BYTES + 128);   // [128] row exponents (split warps only)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nkb = (p.K + TK - 1) / TK;
    const int tiles_n = (p.N + TN - 1) / TN;
    const int num_tiles = tiles_n * ((p.M + TM - 1) / TM);

    if (threadIdx.x == 0) {
        if (smem_u32(smem) & 1023u) __trap();
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&split[s], 256);
            mbar_init(&empty[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&acc_full[b], 1);
            mbar_init(&acc_empty[b], 256);
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
        if (lane == 0) {
            uint32_t it = 0, tcount = 0;
            TRACE(0, 14);
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tcount) {
                const int m0 = (tile / tiles_n) * TM, n0 = (tile % tiles_n) * TN;
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1;
                    mbar_wait(&empty[s], ph ^ 1);
                    if (kb == 0) TRACE(tcount, 0);
                    uint8_t* st = smem + s * STAGE_BYTES;
                    mbar_expect_tx(&full[s], (p.presplit ? 2 * A_H : A_RAW) + 2 * W_H);
                    if (p.presplit) {
                        tma_load_2d(st + A_RAW, &mapA, &full[s], kb * TK, m0);
                        tma_load_2d(st + A_RAW + A_H, &mapAlo, &full[s], kb * TK, m0);
                    } else {
                        tma_load_2d(st, &mapA, &full[s], kb * TK, m0);
                    }
                    tma_load_2d(st + A_RAW + 2 * A_H, &mapWhi, &full[s], kb * TK, n0);
                    tma_load_2d(st + A_RAW + 2 * A_H + W_H, &mapWlo, &full[s], kb * TK, n0);
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            uint32_t it = 0, tcount = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tcount) {
                const uint32_t ab = tcount & 1;                  // accumulator buffer of this tile
                const uint32_t acc = tmem_base + ab * C::ACC_COLS;
                mbar_wait(&acc_empty[ab], ((tcount >> 1) & 1) ^ 1);   // the tile two back has been read out of this buffer
                TRACE(tcount, 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1;
                    mbar_wait(&full[s], ph);
                    if (!p.presplit) mbar_wait(&split[s], ph);
                    if (kb == 0) TRACE(tcount, 2);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t st = smem_u32(smem + s * STAGE_BYTES);
                    const uint64_t d_ahi = umma_desc(st + A_RAW), d_alo = umma_desc(st + A_RAW + A_H);
                    const uint64_t d_whi = umma_desc(st + A_RAW + 2 * A_H), d_wlo = umma_desc(st + A_RAW + 2 * A_H + W_H);
#pragma unroll
                    for (int k = 0; k < TK / 16; ++k) {
                        const uint64_t adv = (uint64_t)((k * 32) >> 4);      // 16 fp16 = 32 bytes along the swizzled row
                        umma_f16(acc, d_ahi + adv, d_whi + adv, IDESC, (kb | k) != 0);
                        umma_f16(acc + CORR, d_alo + adv, d_whi + adv, IDESC, MERGED ? 1u : (uint32_t)((kb | k) != 0));
                        umma_f16(acc + CORR, d_ahi + adv, d_wlo + adv, IDESC, 1u);
                    }
                    umma_commit(&empty[s]);
                }
                umma_commit(&acc_full[ab]);
                TRACE(tcount, 3);
            }
        }
    } else if (warp < 10) {
        // ===================== operand split warps (w2..9) =====================
        const int t = threadIdx.x - 64;     // 0..255
        const mi_epilogue_t& e = p.e;
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < num_tiles && !p.presplit; tile += gridDim.x) {
            const int m0 = (tile / tiles_n) * TM;
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
                asm volatile("bar.sync 1, 256;" ::: "memory");      // everyone is done reading the previous tile's exponents
                if (t < TM) rexp[t] = (int8_t)e8;
                asm volatile("bar.sync 1, 256;" ::: "memory");
            }
            for (int kb = 0; kb < nkb; ++kb, ++it) {
                const int s = it % STAGES;
                const uint32_t ph = (it / STAGES) & 1;
                mbar_wait(&full[s], ph);
                const float4* raw = reinterpret_cast<const float4*>(smem + s * STAGE_BYTES);
                uint8_t* hi = smem + s * STAGE_BYTES + A_RAW;
                uint8_t* lo = hi + A_H;
                // all loads first (the stores below may alias them as far as the compiler knows), then convert + store
                float4 v[4];
                float sc[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {                         // 4 float4 per thread
                    const int pidx = i * 256 + t;                     // physical float4 slot in the 128B-swizzled fp32 tile
                    v[i] = raw[pidx];
                    sc[i] = __uint_as_float((uint32_t)(127 - (int)rexp[pidx >> 3]) << 23);      // 2^-e, exact
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int pidx = i * 256 + t;
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
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> async proxy (UMMA)
                mbar_arrive(&split[s]);
#ifdef MI_TC_TRACE
                if (t == 0 && kb == 0) TRACE((tile - (int)blockIdx.x) / (int)gridDim.x, 13);
#endif
            }
        }
    } else {
        // ===================== epilogue warps (w10..17), overlapped with the next tile's main loop =====================
        const int q = warp & 3;                      // TMEM lane quarter this warp may access
        const int hf = (warp - 10) >> 2;             // column half handled by this warp
        constexpr int EP = C::EPITCH;
        float* ebuf = ebuf_all + (warp - 10) * (32 * EP);
        const mi_epilogue_t& e = p.e;
        uint32_t tcount = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tcount) {
            const int m0 = (tile / tiles_n) * TM, n0 = (tile % tiles_n) * TN;
            const uint32_t ab = tcount & 1;
            const uint32_t acc = tmem_base + ab * C::ACC_COLS;
            // ---- epilogue: TMEM -> registers -> per-warp smem transpose -> coalesced global traffic
            const int mrow = m0 + q * 32 + lane;                 // the row this lane owns in TMEM
            int e8 = 0;                                          // same power-of-two row exponent as the split warps
            if (e.a_amax && mrow < p.M) {
                const int ex = (int)((__float_as_uint(__ldg(e.a_amax + mrow)) >> 23) & 0xff) - 127;
                e8 = max(-100, min(ex - 14, 100));
            }
            const float rowsc = e.alpha * __uint_as_float((uint32_t)(127 + e8) << 23);   // alpha * 2^e
            mbar_wait(&acc_full[ab], (tcount >> 1) & 1);
            if (threadIdx.x == 320) TRACE(tcount, 4);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            int i1 = 0, i2 = 0, i3 = 0;
            if (EPI & 1) {
                const bool mrow_ok = mrow < p.M;
                i1 = (mrow_ok && e.g1) ? (e.g1_idx ? __ldg(e.g1_idx + mrow) : mrow) : 0;
                i2 = (mrow_ok && e.g2) ? (e.g2_idx ? __ldg(e.g2_idx + mrow) : mrow) : 0;
                i3 = (mrow_ok && e.g3) ? (e.g3_idx ? __ldg(e.g3_idx + mrow) : mrow) : 0;
            }
            constexpr int CH = TN / 64;                          // 32-column chunks per warp
            float rowmax[8];                                     // running max |C| of the 8 rows this lane touches
#pragma unroll
            for (int u = 0; u < 8; ++u) rowmax[u] = 0.f;
#pragma unroll 1
            for (int cc = 0; cc < CH; ++cc) {
                const int c = hf * CH + cc;
                const int nb = n0 + c * 32;
                uint32_t v[32], w[32];
                const uint32_t taddr = acc + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32);
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                      "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                      "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                      "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                    : "r"(taddr));
                if (!MERGED) {
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                    : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]),
                      "=r"(w[8]), "=r"(w[9]), "=r"(w[10]), "=r"(w[11]), "=r"(w[12]), "=r"(w[13]), "=r"(w[14]), "=r"(w[15]),
                      "=r"(w[16]), "=r"(w[17]), "=r"(w[18]), "=r"(w[19]), "=r"(w[20]), "=r"(w[21]), "=r"(w[22]), "=r"(w[23]),
                      "=r"(w[24]), "=r"(w[25]), "=r"(w[26]), "=r"(w[27]), "=r"(w[28]), "=r"(w[29]), "=r"(w[30]), "=r"(w[31])
                    : "r"(taddr + CORR));
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) w[j] = 0u;
                }
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (threadIdx.x == 320 && cc < 4) TRACE(tcount, 5 + cc);
                if (cc == CH - 1) {                      // all of this warp's TMEM reads are done: release the accumulators
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    mbar_arrive(&acc_empty[ab]);
                }
                if (nb >= p.N) continue;                 // warp-uniform
#pragma unroll
#pragma unroll
                for (int j = 0; j < 32; j += 2)          // STS.64, bank = (2*lane + j) % 32: conflict-free
                    *reinterpret_cast<float2*>(ebuf + lane * EP + j) = make_float2(
                        rowsc * fmaf(__uint_as_float(w[j]), LO_UNSCALE, __uint_as_float(v[j])),
                        rowsc * fmaf(__uint_as_float(w[j + 1]), LO_UNSCALE, __uint_as_float(v[j + 1])));
                __syncwarp();
                if (threadIdx.x == 320 && cc == 0) TRACE(tcount, 16);
                const int col4 = (lane & 7) * 4;
                const int n = nb + col4;
                if (p.c_vec && nb + 32 <= p.N) {
                    // ---- fast path: whole chunk in range, 16-byte accesses.  Software-pipelined in two batches of four
                    // row groups: all gather / residual loads of a batch are issued before any arithmetic or store, so
                    // their L2 round trips overlap (the compiler will not move loads across the stores by itself).
                    float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (e.bias) bias4 = __ldg(reinterpret_cast<const float4*>(e.bias + n));
#pragma unroll
                    for (int hb = 0; hb < 2; ++hb) {
                        float4 ga[4], gb[4], gc[4], gr[4];
                        int mm[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int rr = (hb * 4 + u) * 4 + (lane >> 3);
                            const int m = m0 + q * 32 + rr;
                            mm[u] = m;
                            const bool ok = m < p.M;
                            ga[u] = gb[u] = gc[u] = gr[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (EPI & 1) {
                                const int r1 = __shfl_sync(0xffffffffu, i1, rr), r2 = __shfl_sync(0xffffffffu, i2, rr),
                                          r3 = __shfl_sync(0xffffffffu, i3, rr);
                                if (ok && e.g1) ga[u] = __ldg(reinterpret_cast<const float4*>(e.g1 + (long long)r1 * e.g1_ld + n));
                                if (ok && e.g2) gb[u] = __ldg(reinterpret_cast<const float4*>(e.g2 + (long long)r2 * e.g2_ld + n));
                                if (ok && e.g3) gc[u] = __ldg(reinterpret_cast<const float4*>(e.g3 + (long long)r3 * e.g3_ld + n));
                            }
                            if (ok && e.resid) gr[u] = __ldg(reinterpret_cast<const float4*>(e.resid + (long long)m * e.resid_ld + n));
                        }
                        if (threadIdx.x == 320 && cc == 0) TRACE(tcount, 17 + 2 * hb);
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int rr = (hb * 4 + u) * 4 + (lane >> 3);
                            const int m = mm[u];
                            if (m >= p.M) continue;
                            const float2 xa = *reinterpret_cast<const float2*>(ebuf + rr * EP + col4);
                            const float2 xb = *reinterpret_cast<const float2*>(ebuf + rr * EP + col4 + 2);
                            float x[4] = {xa.x + bias4.x, xa.y + bias4.y, xb.x + bias4.z, xb.y + bias4.w};
                            if (EPI & 1) {
                                x[0] += ga[u].x + gb[u].x + gc[u].x; x[1] += ga[u].y + gb[u].y + gc[u].y;
                                x[2] += ga[u].z + gb[u].z + gc[u].z; x[3] += ga[u].w + gb[u].w + gc[u].w;
                            }
                            if (EPI & 2) *reinterpret_cast<float4*>(e.z_out + (long long)m * e.z_ld + n) = make_float4(x[0], x[1], x[2], x[3]);
                            if (e.act == MI_ACT_SILU) {
#pragma unroll
                                for (int v4 = 0; v4 < 4; ++v4) x[v4] = silu_fast(x[v4]);
                            }
                            x[0] += gr[u].x; x[1] += gr[u].y; x[2] += gr[u].z; x[3] += gr[u].w;
#ifndef MI_TC_NOSTORE
                            *reinterpret_cast<float4*>(p.C + (long long)m * p.ldc + n) = make_float4(x[0], x[1], x[2], x[3]);
#else
                            if (x[0] == 1.2345f) p.C[0] = x[1] + x[2] + x[3];
#endif
                            rowmax[hb * 4 + u] = fmaxf(rowmax[hb * 4 + u], fmaxf(fmaxf(fabsf(x[0]), fabsf(x[1])), fmaxf(fabsf(x[2]), fabsf(x[3]))));
                        }
                        if (threadIdx.x == 320 && cc == 0) TRACE(tcount, 18 + 2 * hb);
                    }
                } else {
                    // ---- generic path (ragged N or unaligned rows): scalar, bounds-checked
#pragma unroll 1
                    for (int rr0 = 0; rr0 < 32; rr0 += 4) {
                        const int rr = rr0 + (lane >> 3);
                        const int m = m0 + q * 32 + rr;
                        int r1 = 0, r2 = 0, r3 = 0;
                        if (EPI & 1) {
                            r1 = __shfl_sync(0xffffffffu, i1, rr);
                            r2 = __shfl_sync(0xffffffffu, i2, rr);
                            r3 = __shfl_sync(0xffffffffu, i3, rr);
                        }
                        float rmax = 0.f;
                        if (m < p.M) {
                            float* crow = p.C + (long long)m * p.ldc + n;
                            for (int u = 0; u < 4; ++u) {
                                if (n + u >= p.N) continue;
                                float y = ebuf[rr * EP + col4 + u];
                                if (e.bias) y += __ldg(e.bias + n + u);
                                if (EPI & 1) {
                                    if (e.g1) y += __ldg(e.g1 + (long long)r1 * e.g1_ld + n + u);
                                    if (e.g2) y += __ldg(e.g2 + (long long)r2 * e.g2_ld + n + u);
                                    if (e.g3) y += __ldg(e.g3 + (long long)r3 * e.g3_ld + n + u);
                                }
                                if (EPI & 2) e.z_out[(long long)m * e.z_ld + n + u] = y;
                                if (e.act == MI_ACT_SILU) y = silu_fast(y);
                                if (e.resid) y += __ldg(e.resid + (long long)m * e.resid_ld + n + u);
                                crow[u] = y;
                                rmax = fmaxf(rmax, fabsf(y));
                            }
                        }
                        rowmax[rr0 >> 2] = fmaxf(rowmax[rr0 >> 2], rmax);
                    }
                }
                __syncwarp();
                if (threadIdx.x == 320 && cc < 4) TRACE(tcount, 9 + cc);
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


template <int STAGES, int TN, int EPI, int MERGED>
int launch_tc(int M, int N, int K, const void* A, const void* A_lo, int lda, const void* W_hi, const void* W_lo, int ldw,
              cudaStream_t s, const TcParams& p) {
    using C = Cfg<STAGES, TN, MERGED>;
    static bool attr = false;
    int rc;
    CUtensorMap mA, mAl, mWh, mWl;
    if ((rc = make_map(&mA, A, M, K, lda, TM, p.presplit != 0)) != MI_OK) return rc;
    if ((rc = make_map(&mAl, p.presplit ? A_lo : A, M, K, lda, TM, p.presplit != 0)) != MI_OK) return rc;
    if ((rc = make_map(&mWh, W_hi, N, K, ldw, TN, true)) != MI_OK) return rc;
    if ((rc = make_map(&mWl, W_lo, N, K, ldw, TN, true)) != MI_OK) return rc;
    if (!attr) {
        MI_CUDA(cudaFuncSetAttribute(tc_gemm_kernel<STAGES, TN, EPI, MERGED>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
        attr = true;
    }
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        MI_CUDA(cudaGetDevice(&dev));
        MI_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    }
    const long long tiles = (long long)mi_div_up(N, TN) * mi_div_up(M, TM);
    const int grid = (int)(tiles < sms ? tiles : sms);        // persistent: one CTA per SM
    tc_gemm_kernel<STAGES, TN, EPI, MERGED><<<grid, TC_THREADS, C::SMEM_BYTES, s>>>(mA, mAl, mWh, mWl, p);
    MI_CHECK_LAUNCH();
    return MI_OK;
}

}  // namespace

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

static int tc_gemm_impl(int M, int N, int K, const void* A, const void* A_lo, int lda, const void* W_hi, const void* W_lo,
                        int ldw, float* C, int ldc, const mi_epilogue_t* epi, int flags, mi_stream_t stream) {
    const bool merged = (flags & MI_TC_MERGED) != 0;
    MI_CHECK_ARG(M >= 0 && N >= 0 && K > 0, "bad dimension");
    if (M == 0 || N == 0) return MI_OK;
    MI_CHECK_ARG(A && W_hi && W_lo && C, "null operand");
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
    MI_CHECK_ARG(p.e.splitk <= 1, "split-K is not available on the tensor-core path");
    MI_CHECK_ARG(p.e.act != MI_ACT_DSILU && p.e.beta == 0.f, "the tensor-core path has no DSILU / accumulate epilogue (use mi_sgemm)");
    bool cv = (ldc % 4 == 0) && mi_host_aligned16(C);
    const mi_epilogue_t& e = p.e;
    if (e.bias) cv = cv && mi_host_aligned16(e.bias);
    if (e.g1) cv = cv && (e.g1_ld % 4 == 0) && mi_host_aligned16(e.g1);
    if (e.g2) cv = cv && (e.g2_ld % 4 == 0) && mi_host_aligned16(e.g2);
    if (e.g3) cv = cv && (e.g3_ld % 4 == 0) && mi_host_aligned16(e.g3);
    if (e.z_out) cv = cv && (e.z_ld % 4 == 0) && mi_host_aligned16(e.z_out);
    if (e.z_in) cv = cv && (e.zin_ld % 4 == 0) && mi_host_aligned16(e.z_in);
    if (e.resid) cv = cv && (e.resid_ld % 4 == 0) && mi_host_aligned16(e.resid);
    p.c_vec = cv;
    p.presplit = A_lo != nullptr;
    if (p.presplit) MI_CHECK_ARG(p.e.a_amax == nullptr, "pre-split A carries no row rescaling");
    if (merged) MI_CHECK_ARG(p.presplit || p.e.a_amax != nullptr, "the merged format needs the row maxima of A (epi->a_amax)");
    // Column-tile width: 128 (two double-buffered {main, correction} accumulator pairs fill the 512 TMEM columns);
    // 64 only for narrow outputs.
    int tn = (N <= 64) ? 64 : 128;
    const char* force = getenv("MI_TC_TN");
    if (force) tn = atoi(force) <= 64 ? 64 : 128;
    cudaStream_t s = (cudaStream_t)stream;
    const int epi_mode = ((p.e.g1 || p.e.g2 || p.e.g3) ? 1 : 0) | (p.e.z_out ? 2 : 0);
#define MI_TC_CASE(ST, TNV, MG)                                                                          \
    switch (epi_mode) {                                                                                  \
        case 0: return launch_tc<ST, TNV, 0, MG>(M, N, K, A, A_lo, lda, W_hi, W_lo, ldw, s, p);          \
        case 1: return launch_tc<ST, TNV, 1, MG>(M, N, K, A, A_lo, lda, W_hi, W_lo, ldw, s, p);          \
        case 2: return launch_tc<ST, TNV, 2, MG>(M, N, K, A, A_lo, lda, W_hi, W_lo, ldw, s, p);          \
        default: return launch_tc<ST, TNV, 3, MG>(M, N, K, A, A_lo, lda, W_hi, W_lo, ldw, s, p);         \
    }
    if (merged) { MI_TC_CASE(3, 256, 1) }
    if (tn == 128) { MI_TC_CASE(4, 128, 0) }
    MI_TC_CASE(4, 64, 0)
#undef MI_TC_CASE
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
