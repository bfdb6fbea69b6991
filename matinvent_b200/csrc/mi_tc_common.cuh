// Device helpers shared by the tensor-core kernels (mi_tc.cu, mi_node.cu): mbarrier / TMA / tcgen05 wrappers, the fp16
// operand split, fast SiLU.  sm_100a only.
#pragma once
#include <cuda.h>
#include <stdlib.h>
#include <cuda_fp16.h>
#include <cudaTypedefs.h>

#include "mi_common.cuh"

namespace mi_tc {

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
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
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
// One lane of a converged warp.  The single-thread roles (TMA producer, MMA issuer) run their loops on ALL lanes and
// predicate only the issuing instructions with this: tcgen05.mma / cp.async.bulk.tensor take their descriptors, addresses
// and coordinates from UNIFORM registers, and inside an `if (lane == 0)` region the compiler cannot prove them uniform — it
// wraps every such instruction in an ELECT / R2UR.BROADCAST x5 / BRA.U.ANY loop, ~90 cycles per MMA and ~170 per TMA
// operation (profiles/r2m_node_pair_experiment.txt).  In warp-uniform control flow the operands stay in uniform registers.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "elect.sync _|p, 0xffffffff;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(pred));
    return pred != 0;
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

// Programmatic dependent launch: a kernel launched with the stream-serialisation attribute starts while its predecessor
// still runs; pdl_wait() returns once the predecessor has completed and its writes are visible, pdl_launch() lets the
// successor start.  Every kernel here does its set-up (barriers, TMEM, tensor-map prefetch) first, then waits, then
// triggers: at most two kernels overlap, and nothing before the wait touches global memory the predecessor works on.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// 16-column variants (rolled epilogue loops) and the store back to TMEM
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]),
          "f"(v[8]), "f"(v[9]), "f"(v[10]), "f"(v[11]), "f"(v[12]), "f"(v[13]), "f"(v[14]), "f"(v[15])
        : "memory");
}

}  // namespace mi_tc

// host side (defined in mi_tc.cu): 2-D row-major [rows, cols] tensor map, box = [box_rows, 32 columns], zero OOB fill;
// fp32 operands use 128-byte swizzle rows, fp16 operands 64-byte rows
int mi_tc_get_encode();
// launch attribute for programmatic dependent launch: OFF unless MI_PDL=1 — measured twice (round 1 on the single-CTA
// kernels, round 2 on the CTA-pair / cluster kernels: 2 617 vs 2 636 us per reverse step at 256 crystals, 1 717 vs 1 710
// at 128, inside the run-to-run spread): the graph replay already leaves ~1 us between kernels and a successor's CTAs
// cannot become resident before the predecessor's leave (shared memory)
inline bool mi_pdl_enabled() {
    static int on = -1;
    if (on < 0) {
        const char* e = getenv("MI_PDL");
        on = (e && e[0] == '1') ? 1 : 0;
    }
    return on != 0;
}
int mi_tc_make_map(CUtensorMap* map, const void* base, long long rows, long long cols, long long ld, int box_rows, bool half);
int mi_tc_make_map_h16(CUtensorMap* map, const void* base, long long rows, long long cols, long long ld, int box_rows);
// one 3-D map over an fp16 (hi, lo) operand pair (lo follows hi by a multiple of 16 bytes): box = [2, box_rows, 32 columns]
int mi_tc_make_map_pair(CUtensorMap* map, const void* hi, const void* lo, long long rows, long long cols, long long ld, int box_rows);
