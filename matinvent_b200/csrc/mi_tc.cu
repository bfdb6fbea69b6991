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

#include "mi_common.cuh"

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
// 16-byte read-only load that does not allocate in L1: the gathered rows are 4 KB apart (they thrash the L1 sets) and
// every value is used once per CTA; reuse between CTAs is served by L2.
__device__ __forceinline__ float4 ldg_stream4(const float* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ float silu_fast(float x) { return __fdividef(x, 1.0f + __expf(-x)); }
// d/dx [x sigmoid(x)] = s (1 + x (1 - s)), same approximations
__device__ __forceinline__ float dsilu_fast(float x) {
    const float s = __fdividef(1.0f, 1.0f + __expf(-x));
    return s * fmaf(x, 1.0f - s, 1.0f);
}
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
    int ksplit;   // > 1: K is cut into ksplit parts handled by different work units; results are ADDED to C with 16-byte
                  // reductions (weight gradients: few output tiles, tens of thousands of reduction rows)
};

// tcgen05.ld of a 32-lane x 32-column fp32 block (lane = row, registers = consecutive columns); completion is
// only guaranteed after tcgen05.wait::ld.
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
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
        if (lane == 0) {
            TRACE(0, 14);
            auto issue_raw = [&](uint32_t idx) {
                const int tl = (int)(idx / (uint32_t)nkb), kb = (int)(idx % (uint32_t)nkb);
                int m0, n0, kb0;
                tile_origin4(tl, m0, n0, kb0);
                const int r = (int)(idx % (uint32_t)RR);
                mbar_wait(&raw_empty[r], ((idx / (uint32_t)RR) & 1) ^ 1);
                mbar_expect_tx(&raw_full[r], A_RAW);
                tma_load_2d(raw_ring + r * A_RAW, &mapA, &raw_full[r], (kb0 + kb) * TK, m0);
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
                if (kb == 0) TRACE(tl, 0);
                uint8_t* st = op_ring + s * OPB;
                mbar_expect_tx(&w_full[s], (PRESPLIT ? 2 * A_H : 0) + 2 * W_H);
                if (PRESPLIT) {
                    tma_load_2d(st, &mapA, &w_full[s], kc, m0);
                    tma_load_2d(st + A_H, &mapAlo, &w_full[s], kc, m0);
                }
                tma_load_2d(st + 2 * A_H, &mapWhi, &w_full[s], kc, n0);
                tma_load_2d(st + 2 * A_H + W_H, &mapWlo, &w_full[s], kc, n0);
                if (!PRESPLIT && a_it < total) issue_raw(a_it++);
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            uint32_t it = 0;
            for (int tl = 0; tl < my_tiles; ++tl) {
                const uint32_t ab = (uint32_t)tl & 1;            // accumulator buffer of this tile
                const uint32_t acc = tmem_base + ab * C::ACC_COLS;
                mbar_wait(&acc_empty[ab], (((uint32_t)tl >> 1) & 1) ^ 1);   // the tile two back has been read out of this buffer
                TRACE(tl, 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int s = (int)(it % (uint32_t)S);
                    const uint32_t ph = (it / (uint32_t)S) & 1;
                    mbar_wait(&w_full[s], ph);
                    if (!PRESPLIT) mbar_wait(&a_ready[s], ph);
                    if (kb == 0) TRACE(tl, 2);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t st = smem_u32(op_ring + s * OPB);
                    const uint64_t d_ahi = umma_desc(st), d_alo = umma_desc(st + A_H);
                    const uint64_t d_whi = umma_desc(st + 2 * A_H), d_wlo = umma_desc(st + 2 * A_H + W_H);
#pragma unroll
                    for (int k = 0; k < TK / 16; ++k) {
                        const uint64_t adv = (uint64_t)((k * 32) >> 4);      // 16 fp16 = 32 bytes along the swizzled row
                        umma_f16(acc, d_ahi + adv, d_whi + adv, IDESC, (kb | k) != 0);
                        umma_f16(acc + CORR, d_alo + adv, d_whi + adv, IDESC, MERGED ? 1u : (uint32_t)((kb | k) != 0));
                        umma_f16(acc + CORR, d_ahi + adv, d_wlo + adv, IDESC, 1u);
                    }
                    umma_commit(&op_empty[s]);
                }
                umma_commit(&acc_full[ab]);
                TRACE(tl, 3);
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
                            if (p.ksplit > 1)
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
    p.ksplit = p.e.splitk > 1 ? p.e.splitk : 1;
    if (p.ksplit > 1)
        MI_CHECK_ARG(!p.e.bias && !p.e.g1 && !p.e.g2 && !p.e.g3 && !p.e.z_out && !p.e.resid && p.e.act == MI_ACT_NONE &&
                     !p.e.amax_out, "split-K adds alpha * A W^T to C: it takes a plain epilogue (row / column scales only)");
    MI_CHECK_ARG(p.e.beta == 0.f, "the tensor-core path has no accumulate epilogue (pass C as resid, or use mi_sgemm)");
    if (p.e.act == MI_ACT_DSILU)
        MI_CHECK_ARG(p.e.z_in && !p.e.g1 && !p.e.g2 && !p.e.g3 && !p.e.z_out && !p.e.resid,
                     "MI_ACT_DSILU on the tensor-core path needs z_in and a plain epilogue otherwise");
    bool cv = (ldc % 4 == 0) && mi_host_aligned16(C);
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
    if (p.presplit) MI_CHECK_ARG(p.e.a_amax == nullptr, "pre-split A carries no row rescaling");
    if (merged) MI_CHECK_ARG(p.presplit || p.e.a_amax != nullptr, "the merged format needs the row maxima of A (epi->a_amax)");
    // Column-tile width: 128 (two double-buffered {main, correction} accumulator pairs fill the 512 TMEM columns);
    // 64 only for narrow outputs.
    int tn = (N <= 64) ? 64 : 128;
    const char* force = getenv("MI_TC_TN");
    if (force) tn = atoi(force) <= 64 ? 64 : 128;
    cudaStream_t s = (cudaStream_t)stream;
    const int epi_mode = p.e.act == MI_ACT_DSILU ? 8 : (((p.e.g1 || p.e.g2) ? 1 : 0) | (p.e.z_out ? 2 : 0) | ((p.e.g3 || p.e.resid) ? 4 : 0));
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
