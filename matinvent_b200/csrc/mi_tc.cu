// Tensor-core path for the dense blocks: FP32-grade GEMM by 3xTF32 split precision on tcgen05.
//
//   C[M,N] = epilogue( A[M,K] . W[N,K]^T ),  A, W fp32 row-major (K contiguous), fp32 accumulate in TMEM.
//
// 1e-4 parity after 2 000 chained score-network evaluations rules out plain TF32 (10-bit mantissa), so every
// product is formed as  a.w ~= a_hi.w_hi + a_lo.w_hi + a_hi.w_lo  with x_hi = tf32(x) (truncation),
// x_lo = tf32(x - x_hi): 3 tcgen05.mma per k-slice, ~2^-22 relative product error (FP32 FFMA: 2^-24).
//   * W_hi / W_lo are split once per weight update (mi_tf32_split) and streamed by TMA;
//   * A is streamed by TMA as raw fp32 and split in shared memory by the epilogue warpgroup while the
//     previous stage's MMAs run (written back as exact TF32 values, so the tensor core's own input
//     rounding mode is irrelevant);
//   * accumulator: 128 lanes x 256 columns fp32 in TMEM, read back with tcgen05.ld for the fused
//     epilogue (bias + up to 3 row gathers + pre-activation store + SiLU + residual), same contract as mi_sgemm.
//
// CTA = 128x256 output tile, 6 warps: w0 TMA producer, w1 MMA issuer + TMEM owner, w2..5 split + epilogue.
// Shared memory: 2 stages x (A_hi 16K | A_lo 16K | W_hi 32K | W_lo 32K) = 192 KB, 128B-swizzled K-major tiles
// (TMA SWIZZLE_128B <-> UMMA SWIZZLE_128B descriptors), BK = 32 floats = one 128-byte swizzle row.
#include <cuda.h>
#include <stdlib.h>
#include <cudaTypedefs.h>

#include "mi_common.cuh"

namespace {

constexpr int TM = 128;                              // CTA tile rows; columns TN in {256, 128, 64} (template)
constexpr int TC_THREADS = 320;   // w0 TMA, w1 MMA, w2..5 split + epilogue, w6..9 epilogue only
// Pipeline shape: TK floats per k-block (32 -> 128-byte swizzle rows, 16 -> 64-byte), STAGES stages;
// both variants use 192 KB: {32, 2} = 2 x 96 KB, {16, 4} = 4 x 48 KB (deeper prefetch, same bytes per flop).
template <int TK, int STAGES, int TN>
struct Cfg {
    static constexpr int A_BYTES = TM * TK * 4;
    static constexpr int W_BYTES = TN * TK * 4;
    static constexpr uint32_t TMEM_COLS = 2 * TN;                 // main + correction accumulators
    // tcgen05 instruction descriptor, kind::tf32: D=f32 (bits 4-5 = 1), A=B=tf32 (bits 7-9, 10-12 = 2),
    // both K-major (bits 15,16 = 0), N>>3 at bits 17-22, M>>4 at bits 24-28.
    static constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
    static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * W_BYTES;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
    static constexpr uint64_t SBO = (8 * TK * 4) >> 4;            // 8-row group pitch, 16-byte units
    static constexpr uint64_t LAYOUT = (TK == 32) ? 2 : 4;        // UMMA SWIZZLE_128B / SWIZZLE_64B
};
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
// K-major, 128B-swizzled operand tile: 8-row groups of 1024 B (SBO), LBO unused (=1), version 1, layout 2.
template <class C>
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= C::SBO << 32;
    d |= (uint64_t)1 << 46;
    d |= C::LAYOUT << 61;
    return d;
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
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
__device__ __forceinline__ float tf32_trunc(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

struct TcParams {
    int M, N, K;
    float* C; int ldc;
    mi_epilogue_t e;
    int c_vec;
};

template <int TK, int STAGES, int TN>
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapWhi,
               const __grid_constant__ CUtensorMap mapWlo, const TcParams p) {
    using C = Cfg<TK, STAGES, TN>;
    constexpr uint32_t TMEM_COLS = C::TMEM_COLS, IDESC = C::IDESC;
    constexpr int A_BYTES = C::A_BYTES, W_BYTES = C::W_BYTES, STAGE_BYTES = C::STAGE_BYTES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    uint64_t* full = bars;                 // [STAGES] TMA landed
    uint64_t* split = bars + STAGES;       // [STAGES] A_hi/A_lo ready
    uint64_t* empty = bars + 2 * STAGES;   // [STAGES] MMAs done with the stage
    uint64_t* acc_full = bars + 3 * STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * STAGES + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
    const int nkb = (p.K + TK - 1) / TK;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&split[s], 128);
            mbar_init(&empty[s], 1);
        }
        mbar_init(acc_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
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
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % STAGES;
                const uint32_t ph = (kb / STAGES) & 1;
                mbar_wait(&empty[s], ph ^ 1);
                uint8_t* st = smem + s * STAGE_BYTES;
                mbar_expect_tx(&full[s], A_BYTES + 2 * W_BYTES);
                tma_load_2d(st, &mapA, &full[s], kb * TK, m0);
                tma_load_2d(st + 2 * A_BYTES, &mapWhi, &full[s], kb * TK, n0);
                tma_load_2d(st + 2 * A_BYTES + W_BYTES, &mapWlo, &full[s], kb * TK, n0);
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % STAGES;
                const uint32_t ph = (kb / STAGES) & 1;
                mbar_wait(&full[s], ph);
                mbar_wait(&split[s], ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t st = smem_u32(smem + s * STAGE_BYTES);
                const uint64_t d_ahi = umma_desc<C>(st), d_alo = umma_desc<C>(st + A_BYTES);
                const uint64_t d_whi = umma_desc<C>(st + 2 * A_BYTES), d_wlo = umma_desc<C>(st + 2 * A_BYTES + W_BYTES);
#pragma unroll
                for (int k = 0; k < TK / 8; ++k) {
                    const uint64_t adv = (uint64_t)((k * 32) >> 4);      // 8 tf32 = 32 bytes along the swizzle row
                    // the tensor core truncates when it adds into the accumulator, so the error grows with the
                    // number of accumulating instructions: keep the 2^-11-sized correction terms out of the
                    // main accumulator (their truncation errors are 2^-11 smaller in their own accumulator)
                    umma_tf32(tmem_base, d_ahi + adv, d_whi + adv, IDESC, (kb | k) != 0);
                    umma_tf32(tmem_base + TN, d_alo + adv, d_whi + adv, IDESC, (kb | k) != 0);
                    umma_tf32(tmem_base + TN, d_ahi + adv, d_wlo + adv, IDESC, 1u);
                }
                umma_commit(&empty[s]);
            }
            umma_commit(acc_full);
        }
    } else {
        // ===================== split warpgroup (then epilogue) =====================
        const int t = threadIdx.x - 64;     // 0..127 for the split warps
        for (int kb = 0; kb < nkb && warp < 6; ++kb) {
            const int s = kb % STAGES;
            const uint32_t ph = (kb / STAGES) & 1;
            mbar_wait(&full[s], ph);
            float4* hi = reinterpret_cast<float4*>(smem + s * STAGE_BYTES);
            float4* lo = reinterpret_cast<float4*>(smem + s * STAGE_BYTES + A_BYTES);
#pragma unroll
            for (int i = 0; i < (TM * TK / 4) / 128; ++i) {       // 8 float4 per thread, elementwise => swizzle-agnostic
                const int idx = i * 128 + t;
                float4 v = hi[idx];
                float4 h = make_float4(tf32_trunc(v.x), tf32_trunc(v.y), tf32_trunc(v.z), tf32_trunc(v.w));
                float4 l = make_float4(tf32_trunc(v.x - h.x), tf32_trunc(v.y - h.y), tf32_trunc(v.z - h.z), tf32_trunc(v.w - h.w));
                hi[idx] = h;
                lo[idx] = l;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> async proxy (UMMA)
            mbar_arrive(&split[s]);
        }
        // ---- epilogue: TMEM -> registers -> global
        mbar_wait(acc_full, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int q = warp & 3;                      // TMEM lane quarter this warp may access
        const int r = q * 32 + lane;
        const int m = m0 + r;
        const mi_epilogue_t& e = p.e;
        const bool row_ok = m < p.M;
        const float* g1r = (row_ok && e.g1) ? e.g1 + (long long)(e.g1_idx ? __ldg(e.g1_idx + m) : m) * e.g1_ld : nullptr;
        const float* g2r = (row_ok && e.g2) ? e.g2 + (long long)(e.g2_idx ? __ldg(e.g2_idx + m) : m) * e.g2_ld : nullptr;
        const float* g3r = (row_ok && e.g3) ? e.g3 + (long long)(e.g3_idx ? __ldg(e.g3_idx + m) : m) * e.g3_ld : nullptr;
        const int hf = (warp - 2) >> 2;             // column half handled by this warp
        for (int c = hf * (TN / 64); c < (hf + 1) * (TN / 64); ++c) {   // TN/32 column chunks, half per warp set
            const int nb = n0 + c * 32;
            if (nb >= p.N) break;                    // warp-uniform
            uint32_t v[32], w[32];
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32);
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]),
                  "=r"(w[8]), "=r"(w[9]), "=r"(w[10]), "=r"(w[11]), "=r"(w[12]), "=r"(w[13]), "=r"(w[14]), "=r"(w[15]),
                  "=r"(w[16]), "=r"(w[17]), "=r"(w[18]), "=r"(w[19]), "=r"(w[20]), "=r"(w[21]), "=r"(w[22]), "=r"(w[23]),
                  "=r"(w[24]), "=r"(w[25]), "=r"(w[26]), "=r"(w[27]), "=r"(w[28]), "=r"(w[29]), "=r"(w[30]), "=r"(w[31])
                : "r"(taddr + (uint32_t)TN));
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                  "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                  "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                  "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(w[j]));
            if (!row_ok) continue;
            float* crow = p.C + (long long)m * p.ldc + nb;
            if (p.c_vec && nb + 32 <= p.N) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    float x[4] = {__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3])};
                    const int n = nb + j;
#pragma unroll
                    for (int u = 0; u < 4; ++u) x[u] *= e.alpha;
                    if (e.bias) { float4 tq = __ldg(reinterpret_cast<const float4*>(e.bias + n)); x[0] += tq.x; x[1] += tq.y; x[2] += tq.z; x[3] += tq.w; }
                    if (g1r) { float4 tq = __ldg(reinterpret_cast<const float4*>(g1r + n)); x[0] += tq.x; x[1] += tq.y; x[2] += tq.z; x[3] += tq.w; }
                    if (g2r) { float4 tq = __ldg(reinterpret_cast<const float4*>(g2r + n)); x[0] += tq.x; x[1] += tq.y; x[2] += tq.z; x[3] += tq.w; }
                    if (g3r) { float4 tq = __ldg(reinterpret_cast<const float4*>(g3r + n)); x[0] += tq.x; x[1] += tq.y; x[2] += tq.z; x[3] += tq.w; }
                    if (e.beta != 0.f) { float4 tq = *reinterpret_cast<const float4*>(crow + j); x[0] += e.beta * tq.x; x[1] += e.beta * tq.y; x[2] += e.beta * tq.z; x[3] += e.beta * tq.w; }
                    if (e.z_out) *reinterpret_cast<float4*>(e.z_out + (long long)m * e.z_ld + n) = make_float4(x[0], x[1], x[2], x[3]);
                    if (e.act == MI_ACT_SILU) {
#pragma unroll
                        for (int u = 0; u < 4; ++u) x[u] = silu_fast(x[u]);
                    } else if (e.act == MI_ACT_DSILU) {
                        float4 tq = __ldg(reinterpret_cast<const float4*>(e.z_in + (long long)m * e.zin_ld + n));
                        x[0] *= mi_dsilu(tq.x); x[1] *= mi_dsilu(tq.y); x[2] *= mi_dsilu(tq.z); x[3] *= mi_dsilu(tq.w);
                    }
                    if (e.resid) { float4 tq = __ldg(reinterpret_cast<const float4*>(e.resid + (long long)m * e.resid_ld + n)); x[0] += tq.x; x[1] += tq.y; x[2] += tq.z; x[3] += tq.w; }
                    *reinterpret_cast<float4*>(crow + j) = make_float4(x[0], x[1], x[2], x[3]);
                }
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int n = nb + j;
                    if (n >= p.N) continue;
                    float x = e.alpha * __uint_as_float(v[j]);
                    if (e.bias) x += __ldg(e.bias + n);
                    if (g1r) x += __ldg(g1r + n);
                    if (g2r) x += __ldg(g2r + n);
                    if (g3r) x += __ldg(g3r + n);
                    if (e.beta != 0.f) x += e.beta * crow[j];
                    if (e.z_out) e.z_out[(long long)m * e.z_ld + n] = x;
                    if (e.act == MI_ACT_SILU) x = silu_fast(x);
                    else if (e.act == MI_ACT_DSILU) x *= mi_dsilu(__ldg(e.z_in + (long long)m * e.zin_ld + n));
                    if (e.resid) x += __ldg(e.resid + (long long)m * e.resid_ld + n);
                    crow[j] = x;
                }
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

__global__ void tf32_split_kernel(const float* __restrict__ w, float* __restrict__ hi, float* __restrict__ lo, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float x = w[i];
    float h = tf32_trunc(x);
    hi[i] = h;
    lo[i] = tf32_trunc(x - h);
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

// 2-D fp32 row-major [rows, cols] (ld elements between rows), box = [box_rows, 32 cols], 128B swizzle, zero OOB fill
int make_map(CUtensorMap* map, const float* base, long long rows, long long cols, long long ld, int box_rows, int tk) {
    cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t gstr[1] = {(cuuint64_t)ld * 4};
    cuuint32_t box[2] = {(cuuint32_t)tk, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstr, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, tk == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        mi_set_error_("cuTensorMapEncodeTiled failed (%d) for [%lld,%lld] ld %lld", (int)r, rows, cols, ld);
        return MI_ERR_CUDA;
    }
    return MI_OK;
}

int g_variant = -1;   // MI_TC_VARIANT env: 0 = {TK 32, 2 stages}, 1 = {TK 16, 4 stages} (default)

template <int TK, int STAGES, int TN>
int launch_tc(int M, int N, int K, const float* A, int lda, const float* W_hi, const float* W_lo, int ldw, cudaStream_t s,
              const TcParams& p) {
    using C = Cfg<TK, STAGES, TN>;
    static bool attr = false;
    int rc;
    CUtensorMap mA, mWh, mWl;
    if ((rc = make_map(&mA, A, M, K, lda, TM, TK)) != MI_OK) return rc;
    if ((rc = make_map(&mWh, W_hi, N, K, ldw, TN, TK)) != MI_OK) return rc;
    if ((rc = make_map(&mWl, W_lo, N, K, ldw, TN, TK)) != MI_OK) return rc;
    if (!attr) {
        MI_CUDA(cudaFuncSetAttribute(tc_gemm_kernel<TK, STAGES, TN>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
        attr = true;
    }
    dim3 grid(mi_div_up(N, TN), mi_div_up(M, TM));
    MI_CHECK_ARG(grid.y <= 65535u, "grid too large");
    tc_gemm_kernel<TK, STAGES, TN><<<grid, TC_THREADS, C::SMEM_BYTES, s>>>(mA, mWh, mWl, p);
    MI_CHECK_LAUNCH();
    return MI_OK;
}

}  // namespace

extern "C" int mi_tf32_split(const float* w, float* hi, float* lo, long long n, mi_stream_t stream) {
    if (n <= 0) return MI_OK;
    MI_CHECK_ARG(w && hi && lo, "null pointer");
    tf32_split_kernel<<<mi_div_up(n, 256), 256, 0, (cudaStream_t)stream>>>(w, hi, lo, n);
    MI_CHECK_LAUNCH();
    return MI_OK;
}

extern "C" int mi_tc_gemm(int M, int N, int K, const float* A, int lda, const float* W_hi, const float* W_lo, int ldw,
                          float* C, int ldc, const mi_epilogue_t* epi, mi_stream_t stream) {
    MI_CHECK_ARG(M >= 0 && N >= 0 && K > 0, "bad dimension");
    if (M == 0 || N == 0) return MI_OK;
    MI_CHECK_ARG(A && W_hi && W_lo && C, "null operand");
    MI_CHECK_ARG(lda >= K && ldw >= K && ldc >= N, "leading dimension too small");
    MI_CHECK_ARG(lda % 4 == 0 && ldw % 4 == 0 && mi_host_aligned16(A) && mi_host_aligned16(W_hi) && mi_host_aligned16(W_lo),
                 "TMA operands need 16-byte aligned rows");
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
    if (p.e.act == MI_ACT_DSILU) MI_CHECK_ARG(p.e.z_in != nullptr, "DSILU epilogue needs z_in");
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
    if (g_variant < 0) {
        const char* v = getenv("MI_TC_VARIANT");
        g_variant = v ? atoi(v) : 1;
    }
    // Column-tile width: small-M node GEMMs and the per-GPU shards of an 8-GPU run leave most SMs idle with
    // 128x256 tiles.
    const long long mt = mi_div_up(M, TM);
    int tn = 256;                                  // measured on B200: 128x128 tiles win below ~half a wave of 128x256 tiles
    if (mt * mi_div_up(N, 256) < 74) tn = 128;
    if (N <= 64) tn = 64;
    else if (N <= 128) tn = 128;
    const char* force = getenv("MI_TC_TN");
    if (force) tn = atoi(force);
    cudaStream_t s = (cudaStream_t)stream;
    if (g_variant == 0 && tn == 256) return launch_tc<32, 2, 256>(M, N, K, A, lda, W_hi, W_lo, ldw, s, p);
    if (tn == 256) return launch_tc<16, 4, 256>(M, N, K, A, lda, W_hi, W_lo, ldw, s, p);
    if (tn == 128) return launch_tc<32, 3, 128>(M, N, K, A, lda, W_hi, W_lo, ldw, s, p);
    return launch_tc<32, 4, 64>(M, N, K, A, lda, W_hi, W_lo, ldw, s, p);
}
