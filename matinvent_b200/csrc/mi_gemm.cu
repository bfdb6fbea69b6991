// FP32 GEMM family with fused gather / activation epilogues (CUDA-core FFMA path).
//
// Replaces the nn.Linear / torch.cat / gather / SiLU chains of models/diffcsp/cspnet.py:45-54,59-82,
// 264-294 (and their autograd backward).  The reverse sampler must stay within 1e-4 of the reference
// after 2 000 chained score-network evaluations, which rules out plain TF32/BF16 tensor-core math
// (SURVEY.md §7 "hard parts"); this kernel is the full-FP32 path used for every dense block.
//
// Tiling: 128x128x16 CTA tile, 256 threads, 8x8 register micro-tile per thread (two 4-wide halves in
// each dimension so shared-memory reads are conflict-free float4s), double-buffered shared memory
// with register-staged global prefetch (one __syncthreads per k-step).
#include "mi_common.cuh"

namespace {

constexpr int BM = 128, BN = 128, BK = 16, NT = 256, PAD = 4;

struct GemmParams {
    int M, N, K;
    const float* A; int lda;
    const float* B; int ldb;
    float* C; int ldc;
    mi_epilogue_t e;
    int a_vec, b_vec, c_vec;  // 16-byte vector access legal for the operand
    int k_per_split;
};

// Load a [rows=128][k=16] tile stored K-contiguous (row stride ld): thread -> 2 x float4 along k.
struct FragKC { float4 v[2]; };
// Load a [k=16][cols=128] tile stored column(M/N)-contiguous: thread -> 2 x float4 along m.
struct FragMC { float4 v[2]; };

__device__ __forceinline__ float4 ldg4_guard(const float* __restrict__ base, long long off, int nvalid,
                                             bool vec) {
    float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
    if (nvalid >= 4 && vec) {
        r = __ldg(reinterpret_cast<const float4*>(base + off));
    } else {
        if (nvalid > 0) r.x = __ldg(base + off);
        if (nvalid > 1) r.y = __ldg(base + off + 1);
        if (nvalid > 2) r.z = __ldg(base + off + 2);
        if (nvalid > 3) r.w = __ldg(base + off + 3);
    }
    return r;
}

// K-contiguous operand: element (r, k) at base[r*ld + k]; tile origin (r0, k0); extents R, Kend.
__device__ __forceinline__ void load_kc(FragKC& f, const float* __restrict__ base, int ld, int r0, int k0,
                                        int R, int Kend, bool vec, int tid) {
    const int kq = (tid & 3) * 4;
    const int rr = tid >> 2;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        int r = r0 + rr + h * 64;
        int k = k0 + kq;
        int nvalid = (r < R) ? min(4, Kend - k) : 0;
        f.v[h] = ldg4_guard(base, (long long)r * ld + k, nvalid, vec);
    }
}
__device__ __forceinline__ void store_kc(const FragKC& f, float (*S)[BM + PAD], int tid) {
    const int kq = (tid & 3) * 4;
    const int rr = tid >> 2;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        int r = rr + h * 64;
        S[kq + 0][r] = f.v[h].x;
        S[kq + 1][r] = f.v[h].y;
        S[kq + 2][r] = f.v[h].z;
        S[kq + 3][r] = f.v[h].w;
    }
}
// Row(M)-contiguous operand: element (r, k) at base[k*ld + r].
__device__ __forceinline__ void load_mc(FragMC& f, const float* __restrict__ base, int ld, int r0, int k0,
                                        int R, int Kend, bool vec, int tid) {
    const int r4 = (tid & 31) * 4;
    const int kk = tid >> 5;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        int k = k0 + kk + h * 8;
        int r = r0 + r4;
        int nvalid = (k < Kend) ? min(4, R - r) : 0;
        f.v[h] = ldg4_guard(base, (long long)k * ld + r, nvalid, vec);
    }
}
__device__ __forceinline__ void store_mc(const FragMC& f, float (*S)[BM + PAD], int tid) {
    const int r4 = (tid & 31) * 4;
    const int kk = tid >> 5;
#pragma unroll
    for (int h = 0; h < 2; ++h) *reinterpret_cast<float4*>(&S[kk + h * 8][r4]) = f.v[h];
}

template <bool A_KC, bool B_KC>
__global__ void __launch_bounds__(NT, 2) sgemm_kernel(const GemmParams p) {
    __shared__ __align__(16) float As[2][BK][BM + PAD];
    __shared__ __align__(16) float Bs[2][BK][BN + PAD];

    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int kbeg = blockIdx.z * p.k_per_split;
    const int kend = min(p.K, kbeg + p.k_per_split);
    const int nk = (kend - kbeg + BK - 1) / BK;

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    FragKC fa_k, fb_k;
    FragMC fa_m, fb_m;

    auto gload = [&](int kt) {
        int k0 = kbeg + kt * BK;
        if (A_KC) load_kc(fa_k, p.A, p.lda, m0, k0, p.M, kend, p.a_vec, tid);
        else      load_mc(fa_m, p.A, p.lda, m0, k0, p.M, kend, p.a_vec, tid);
        if (B_KC) load_kc(fb_k, p.B, p.ldb, n0, k0, p.N, kend, p.b_vec, tid);
        else      load_mc(fb_m, p.B, p.ldb, n0, k0, p.N, kend, p.b_vec, tid);
    };
    auto sstore = [&](int buf) {
        if (A_KC) store_kc(fa_k, As[buf], tid); else store_mc(fa_m, As[buf], tid);
        if (B_KC) store_kc(fb_k, Bs[buf], tid); else store_mc(fb_m, Bs[buf], tid);
    };

    if (nk > 0) {
        gload(0);
        sstore(0);
    }
    __syncthreads();
    int cur = 0;
    for (int kt = 0; kt < nk; ++kt) {
        if (kt + 1 < nk) gload(kt + 1);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float4 a0 = *reinterpret_cast<const float4*>(&As[cur][k][ty * 4]);
            float4 a1 = *reinterpret_cast<const float4*>(&As[cur][k][64 + ty * 4]);
            float4 b0 = *reinterpret_cast<const float4*>(&Bs[cur][k][tx * 4]);
            float4 b1 = *reinterpret_cast<const float4*>(&Bs[cur][k][64 + tx * 4]);
            float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (kt + 1 < nk) sstore(cur ^ 1);
        __syncthreads();
        cur ^= 1;
    }

    // ------------------------------------------------------------------ epilogue
    const mi_epilogue_t& e = p.e;
    const bool atomic = e.splitk > 1;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = m0 + ((i < 4) ? (ty * 4 + i) : (64 + ty * 4 + (i - 4)));
        if (m >= p.M) continue;
        const float* g1r = e.g1 ? e.g1 + (long long)(e.g1_idx ? __ldg(e.g1_idx + m) : m) * e.g1_ld : nullptr;
        const float* g2r = e.g2 ? e.g2 + (long long)(e.g2_idx ? __ldg(e.g2_idx + m) : m) * e.g2_ld : nullptr;
        const float* g3r = e.g3 ? e.g3 + (long long)(e.g3_idx ? __ldg(e.g3_idx + m) : m) * e.g3_ld : nullptr;
        float rmax = 0.f;                   // max |C| of this thread's 8 columns of row m
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int n = n0 + h * 64 + tx * 4;
            if (n >= p.N) continue;
            float v[4] = {acc[i][h * 4 + 0], acc[i][h * 4 + 1], acc[i][h * 4 + 2], acc[i][h * 4 + 3]};
            float* crow = p.C + (long long)m * p.ldc + n;
            if (atomic) {
                if (p.c_vec && n + 3 < p.N) {
                    // one 16-byte reduction instead of four scalar atomics: the weight-gradient GEMMs of the fine-tune
                    // step (K split 11 ways at the reference's batch) were bound by L2 atomic throughput
                    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(crow), "f"(e.alpha * v[0]), "f"(e.alpha * v[1]),
                                 "f"(e.alpha * v[2]), "f"(e.alpha * v[3])
                                 : "memory");
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (n + j < p.N) atomicAdd(crow + j, e.alpha * v[j]);
                }
                continue;
            }
            if (p.c_vec && n + 3 < p.N) {
#pragma unroll
                for (int j = 0; j < 4; ++j) v[j] *= e.alpha;
                if (e.bias) {
                    float4 t = __ldg(reinterpret_cast<const float4*>(e.bias + n));
                    v[0] += t.x; v[1] += t.y; v[2] += t.z; v[3] += t.w;
                }
                if (g1r) { float4 t = __ldg(reinterpret_cast<const float4*>(g1r + n)); v[0] += t.x; v[1] += t.y; v[2] += t.z; v[3] += t.w; }
                if (g2r) { float4 t = __ldg(reinterpret_cast<const float4*>(g2r + n)); v[0] += t.x; v[1] += t.y; v[2] += t.z; v[3] += t.w; }
                if (g3r) { float4 t = __ldg(reinterpret_cast<const float4*>(g3r + n)); v[0] += t.x; v[1] += t.y; v[2] += t.z; v[3] += t.w; }
                if (e.beta != 0.f) {
                    float4 t = *reinterpret_cast<const float4*>(crow);
                    v[0] += e.beta * t.x; v[1] += e.beta * t.y; v[2] += e.beta * t.z; v[3] += e.beta * t.w;
                }
                if (e.z_out)
                    *reinterpret_cast<float4*>(e.z_out + (long long)m * e.z_ld + n) = make_float4(v[0], v[1], v[2], v[3]);
                if (e.act == MI_ACT_SILU) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) v[j] = mi_silu(v[j]);
                } else if (e.act == MI_ACT_DSILU) {
                    float4 t = __ldg(reinterpret_cast<const float4*>(e.z_in + (long long)m * e.zin_ld + n));
                    v[0] *= mi_dsilu(t.x); v[1] *= mi_dsilu(t.y); v[2] *= mi_dsilu(t.z); v[3] *= mi_dsilu(t.w);
                }
                if (e.resid) {
                    float4 t = __ldg(reinterpret_cast<const float4*>(e.resid + (long long)m * e.resid_ld + n));
                    v[0] += t.x; v[1] += t.y; v[2] += t.z; v[3] += t.w;
                }
                *reinterpret_cast<float4*>(crow) = make_float4(v[0], v[1], v[2], v[3]);
                rmax = fmaxf(rmax, fmaxf(fmaxf(fabsf(v[0]), fabsf(v[1])), fmaxf(fabsf(v[2]), fabsf(v[3]))));
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (n + j >= p.N) continue;
                    float x = e.alpha * v[j];
                    if (e.bias) x += __ldg(e.bias + n + j);
                    if (g1r) x += __ldg(g1r + n + j);
                    if (g2r) x += __ldg(g2r + n + j);
                    if (g3r) x += __ldg(g3r + n + j);
                    if (e.beta != 0.f) x += e.beta * crow[j];
                    if (e.z_out) e.z_out[(long long)m * e.z_ld + n + j] = x;
                    if (e.act == MI_ACT_SILU) x = mi_silu(x);
                    else if (e.act == MI_ACT_DSILU) x *= mi_dsilu(__ldg(e.z_in + (long long)m * e.zin_ld + n + j));
                    if (e.resid) x += __ldg(e.resid + (long long)m * e.resid_ld + n + j);
                    crow[j] = x;
                    rmax = fmaxf(rmax, fabsf(x));
                }
            }
        }
        if (e.amax_out && !atomic) {
            // the 16 threads of a half warp hold the 128 columns of row m: one atomic per row and CTA tile
            // (one per float4 serialised on the row's address: 39 us for the 2643 x 512 embedding GEMM)
            const unsigned hm = (threadIdx.x & 16) ? 0xffff0000u : 0x0000ffffu;
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) rmax = fmaxf(rmax, __shfl_xor_sync(hm, rmax, o));
            if ((threadIdx.x & 15) == 0) atomicMax(reinterpret_cast<unsigned*>(e.amax_out + m), __float_as_uint(rmax));
        }
    }
}

}  // namespace

extern "C" int mi_sgemm(int transA, int transB, int M, int N, int K, const float* A, int lda,
                        const float* B, int ldb, float* C, int ldc, const mi_epilogue_t* epi,
                        mi_stream_t stream) {
    MI_CHECK_ARG(M >= 0 && N >= 0 && K >= 0, "negative dimension");
    if (M == 0 || N == 0) return MI_OK;
    MI_CHECK_ARG(A && B && C, "null operand");
    MI_CHECK_ARG(lda >= (transA ? M : K) && ldb >= (transB ? K : N) && ldc >= N, "leading dimension too small");
    GemmParams p;
    p.M = M; p.N = N; p.K = K; p.A = A; p.lda = lda; p.B = B; p.ldb = ldb; p.C = C; p.ldc = ldc;
    if (epi) p.e = *epi;
    else {
        mi_epilogue_t z = {};
        z.alpha = 1.f; z.splitk = 1;
        p.e = z;
    }
    MI_CHECK_ARG(p.e.col_scale == nullptr, "col_scale belongs to the tensor-core path (mi_tc_gemm)");
    MI_CHECK_ARG(p.e.scat_out == nullptr, "the fused scatter epilogue belongs to the tensor-core path (mi_tc_gemm)");
    if (p.e.splitk < 1) p.e.splitk = 1;
    if (p.e.splitk > 1) {
        MI_CHECK_ARG(!p.e.amax_out, "split-K cannot report row maxima");
        MI_CHECK_ARG(!p.e.bias && !p.e.g1 && !p.e.g2 && !p.e.g3 && !p.e.z_out && !p.e.resid &&
                     p.e.act == MI_ACT_NONE && p.e.beta == 1.f, "split-K needs a plain accumulate epilogue");
    }
    if (p.e.act == MI_ACT_DSILU) MI_CHECK_ARG(p.e.z_in != nullptr, "DSILU epilogue needs z_in");
    p.a_vec = (lda % 4 == 0) && mi_host_aligned16(A);
    p.b_vec = (ldb % 4 == 0) && mi_host_aligned16(B);
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
    int splitk = p.e.splitk;
    int kps = ((K + splitk - 1) / splitk + BK - 1) / BK * BK;
    if (kps <= 0) kps = BK;
    splitk = (K + kps - 1) / kps;
    if (splitk < 1) splitk = 1;
    p.k_per_split = kps;
    if (p.e.splitk > 1 && splitk == 1) {
        // degenerate split: still use the atomic accumulate path (beta == 1 semantics)
        p.e.splitk = 2;
    }
    dim3 grid(mi_div_up(N, BN), mi_div_up(M, BM), splitk);
    MI_CHECK_ARG(grid.y <= 65535u && grid.z <= 65535u, "grid too large");
    cudaStream_t s = (cudaStream_t)stream;
    if (!transA && transB) sgemm_kernel<true, true><<<grid, NT, 0, s>>>(p);
    else if (!transA && !transB) sgemm_kernel<true, false><<<grid, NT, 0, s>>>(p);
    else if (transA && !transB) sgemm_kernel<false, false><<<grid, NT, 0, s>>>(p);
    else sgemm_kernel<false, true><<<grid, NT, 0, s>>>(p);
    MI_CHECK_LAUNCH();
    return MI_OK;
}
