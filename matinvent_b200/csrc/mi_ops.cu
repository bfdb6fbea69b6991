// Graph construction, segment reductions, LayerNorm, small per-crystal pieces, reverse-diffusion
// updates, forward noising, losses, Adam and RNG.  All HBM-bound: coalesced 128-bit accesses, one
// pass over the data, grids sized from the problem (every kernel is a few waves at most).
#include <stdarg.h>
#include <string.h>

#include <cuda_fp16.h>

#include "mi_common.cuh"

// ------------------------------------------------------------------------------------ error plumbing
static thread_local char g_err[512] = "";
extern "C" void mi_set_error_(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
extern "C" const char* mi_last_error(void) { return g_err; }
extern "C" int mi_version(void) { return 100; }
extern "C" int mi_device_info(int* sm_count, int* cc_major, int* cc_minor) {
    int dev = 0;
    MI_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp pr;
    MI_CUDA(cudaGetDeviceProperties(&pr, dev));
    if (sm_count) *sm_count = pr.multiProcessorCount;
    if (cc_major) *cc_major = pr.major;
    if (cc_minor) *cc_minor = pr.minor;
    return MI_OK;
}

namespace {

__device__ __forceinline__ int upper_seg(const int* __restrict__ off, int B, int e) {
    // largest b in [0,B) with off[b] <= e
    int lo = 0, hi = B;
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (__ldg(off + mid) <= e) lo = mid; else hi = mid;
    }
    return lo;
}

// ------------------------------------------------------------------------------------ fc edges
__global__ void fc_edges_kernel(const int* __restrict__ node_off, const int* __restrict__ edge_off, int B,
                                int N, int E, int* __restrict__ edge_src, int* __restrict__ edge_dst,
                                int* __restrict__ edge_graph, int* __restrict__ seg_ptr,
                                int* __restrict__ dst_ptr, int* __restrict__ dst_perm,
                                int* __restrict__ node_graph) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < E) {
        int b = upper_seg(edge_off, B, t);
        int n0 = __ldg(node_off + b), n = __ldg(node_off + b + 1) - n0, e0 = __ldg(edge_off + b);
        int le = t - e0;
        int i = le / n, j = le - i * n;
        edge_src[t] = n0 + i;
        edge_dst[t] = n0 + j;
        edge_graph[t] = b;
        if (dst_perm) dst_perm[e0 + j * n + i] = t;
    }
    if (t < N) {
        int b = upper_seg(node_off, B, t);
        int n0 = __ldg(node_off + b), n = __ldg(node_off + b + 1) - n0, e0 = __ldg(edge_off + b);
        int p = e0 + (t - n0) * n;
        seg_ptr[t] = p;
        if (dst_ptr) dst_ptr[t] = p;
        node_graph[t] = b;
    }
    if (t == 0) {
        seg_ptr[N] = E;
        if (dst_ptr) dst_ptr[N] = E;
    }
}

// ------------------------------------------------------------------------------------ Fourier basis
// thread = (edge, coordinate c, 4 consecutive frequencies): four independent sincos chains per thread (ILP) and
// 8/16-byte stores; the frequency block of a coordinate is 4-aligned because F % 4 == 0 on this path.
__global__ void edge_fourier_kernel(const float* __restrict__ x, const int* __restrict__ src,
                                    const int* __restrict__ dst, const float* __restrict__ cell_off,
                                    int E, int F, float* __restrict__ frac_diff, float* __restrict__ phi,
                                    int ld_phi, __half* __restrict__ phi_hi, __half* __restrict__ phi_lo,
                                    float op_scale, float lo_scale) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int F3 = 3 * F, Q = F3 >> 2;
    if (t >= (long long)E * Q) return;
    int e = (int)(t / Q);
    int col = (int)(t - (long long)e * Q) << 2;
    int c = col / F, k = col - c * F;
    int i = __ldg(src + e), j = __ldg(dst + e);
    float d = __fsub_rn(__ldg(x + 3 * j + c), __ldg(x + 3 * i + c));
    if (cell_off) d = __fadd_rn(d, __ldg(cell_off + 3 * (long long)e + c));   // knn: un-wrapped (cspnet.py:252-257)
    else d = mi_mod1(d);                                                      // fc: (x_j - x_i) % 1 (cspnet.py:242)
    if (k == 0 && frac_diff) frac_diff[3 * (long long)e + c] = d;
    float sv[4], cv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        float freq = __fmul_rn((float)(k + u), 6.2831855f);   // fl32(2*pi) * fl32(k)   (cspnet.py:18)
        sincosf(__fmul_rn(d, freq), &sv[u], &cv[u]);
    }
    const long long o = (long long)e * ld_phi + col;
    if (phi) {
        *reinterpret_cast<float4*>(phi + o) = make_float4(sv[0], sv[1], sv[2], sv[3]);
        *reinterpret_cast<float4*>(phi + o + F3) = make_float4(cv[0], cv[1], cv[2], cv[3]);
    }
    if (phi_hi) {      // operand form of mi_tc_gemm_presplit: fp16 head + 2^11-scaled fp16 tail (|Phi| <= 1: no rescaling)
#pragma unroll
        for (int u = 0; u < 4; ++u) { sv[u] *= op_scale; cv[u] *= op_scale; }
        __half2 hs0 = __floats2half2_rn(sv[0], sv[1]), hs1 = __floats2half2_rn(sv[2], sv[3]);
        __half2 hc0 = __floats2half2_rn(cv[0], cv[1]), hc1 = __floats2half2_rn(cv[2], cv[3]);
        float2 fs0 = __half22float2(hs0), fs1 = __half22float2(hs1), fc0 = __half22float2(hc0), fc1 = __half22float2(hc1);
        __half2 ls0 = __floats2half2_rn((sv[0] - fs0.x) * lo_scale, (sv[1] - fs0.y) * lo_scale);
        __half2 ls1 = __floats2half2_rn((sv[2] - fs1.x) * lo_scale, (sv[3] - fs1.y) * lo_scale);
        __half2 lc0 = __floats2half2_rn((cv[0] - fc0.x) * lo_scale, (cv[1] - fc0.y) * lo_scale);
        __half2 lc1 = __floats2half2_rn((cv[2] - fc1.x) * lo_scale, (cv[3] - fc1.y) * lo_scale);
        auto st2 = [](__half* p, __half2 a, __half2 b) {
            uint2 v;
            v.x = *reinterpret_cast<unsigned*>(&a);
            v.y = *reinterpret_cast<unsigned*>(&b);
            *reinterpret_cast<uint2*>(p) = v;
        };
        st2(phi_hi + o, hs0, hs1);
        st2(phi_hi + o + F3, hc0, hc1);
        st2(phi_lo + o, ls0, ls1);
        st2(phi_lo + o + F3, lc0, lc1);
    }
}
// scalar variant for F % 4 != 0 (thread = one (edge, column))
__global__ void edge_fourier_kernel_scalar(const float* __restrict__ x, const int* __restrict__ src,
                                           const int* __restrict__ dst, const float* __restrict__ cell_off,
                                           int E, int F, float* __restrict__ frac_diff, float* __restrict__ phi,
                                           int ld_phi, __half* __restrict__ phi_hi, __half* __restrict__ phi_lo,
                                           float op_scale, float lo_scale) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int F3 = 3 * F;
    if (t >= (long long)E * F3) return;
    int e = (int)(t / F3);
    int col = (int)(t - (long long)e * F3);
    int c = col / F, k = col - c * F;
    int i = __ldg(src + e), j = __ldg(dst + e);
    float d = __fsub_rn(__ldg(x + 3 * j + c), __ldg(x + 3 * i + c));
    if (cell_off) d = __fadd_rn(d, __ldg(cell_off + 3 * (long long)e + c));
    else d = mi_mod1(d);
    if (k == 0 && frac_diff) frac_diff[3 * (long long)e + c] = d;
    float freq = __fmul_rn((float)k, 6.2831855f);
    float s, co;
    sincosf(__fmul_rn(d, freq), &s, &co);
    const long long o = (long long)e * ld_phi;
    if (phi) {
        phi[o + col] = s;
        phi[o + F3 + col] = co;
    }
    if (phi_hi) {
        s *= op_scale;
        co *= op_scale;
        __half hs = __float2half_rn(s), hc = __float2half_rn(co);
        phi_hi[o + col] = hs;
        phi_hi[o + F3 + col] = hc;
        phi_lo[o + col] = __float2half_rn((s - __half2float(hs)) * lo_scale);
        phi_lo[o + F3 + col] = __float2half_rn((co - __half2float(hc)) * lo_scale);
    }
}

// ------------------------------------------------------------------------------------ segment reduce
// Persistent blocks (a multiple of the SM count) walk the segments grid-stride; thread = one float4 column of
// one segment; 8 independent 128-bit streaming loads in flight per thread (rows of a segment are consecutive
// 2 KB lines, so every warp request is fully coalesced).
__device__ __forceinline__ float4 ld_stream(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__global__ void __launch_bounds__(128) segment_reduce_kernel(const float* __restrict__ X, int ldx,
                                                             const int* __restrict__ ptr,
                                                             const int* __restrict__ perm,
                                                             float* __restrict__ out, int ldo, int H4, int S,
                                                             int mean, int accumulate, float* __restrict__ amax_out) {
    const int c = blockIdx.y * blockDim.x + threadIdx.x;
    const bool live = c < H4;
    const long long ld4 = ldx >> 2;
    const float4* base = reinterpret_cast<const float4*>(X) + c;
    __shared__ float wm[4];
    for (int s = blockIdx.x; s < S; s += gridDim.x) {
        float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
        if (live) {
            const int beg = __ldg(ptr + s), end = __ldg(ptr + s + 1);
            float4 a0 = r, a1 = r, a2 = r, a3 = r;
            int k = beg;
            for (; k + 8 <= end; k += 8) {
                long long rr[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) rr[u] = perm ? (long long)__ldg(perm + k + u) : (long long)(k + u);
                float4 v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) v[u] = ld_stream(base + rr[u] * ld4);
#pragma unroll
                for (int u = 0; u < 8; u += 4) {
                    a0.x += v[u].x; a0.y += v[u].y; a0.z += v[u].z; a0.w += v[u].w;
                    a1.x += v[u + 1].x; a1.y += v[u + 1].y; a1.z += v[u + 1].z; a1.w += v[u + 1].w;
                    a2.x += v[u + 2].x; a2.y += v[u + 2].y; a2.z += v[u + 2].z; a2.w += v[u + 2].w;
                    a3.x += v[u + 3].x; a3.y += v[u + 3].y; a3.z += v[u + 3].z; a3.w += v[u + 3].w;
                }
            }
            for (; k + 4 <= end; k += 4) {
                long long r0 = perm ? __ldg(perm + k) : k, r1 = perm ? __ldg(perm + k + 1) : k + 1;
                long long r2 = perm ? __ldg(perm + k + 2) : k + 2, r3 = perm ? __ldg(perm + k + 3) : k + 3;
                float4 v0 = ld_stream(base + r0 * ld4), v1 = ld_stream(base + r1 * ld4);
                float4 v2 = ld_stream(base + r2 * ld4), v3 = ld_stream(base + r3 * ld4);
                a0.x += v0.x; a0.y += v0.y; a0.z += v0.z; a0.w += v0.w;
                a1.x += v1.x; a1.y += v1.y; a1.z += v1.z; a1.w += v1.w;
                a2.x += v2.x; a2.y += v2.y; a2.z += v2.z; a2.w += v2.w;
                a3.x += v3.x; a3.y += v3.y; a3.z += v3.z; a3.w += v3.w;
            }
            for (; k < end; ++k) {
                long long r0 = perm ? __ldg(perm + k) : k;
                float4 v0 = ld_stream(base + r0 * ld4);
                a0.x += v0.x; a0.y += v0.y; a0.z += v0.z; a0.w += v0.w;
            }
            r = make_float4((a0.x + a1.x) + (a2.x + a3.x), (a0.y + a1.y) + (a2.y + a3.y),
                            (a0.z + a1.z) + (a2.z + a3.z), (a0.w + a1.w) + (a2.w + a3.w));
            if (mean) {
                float cnt = (float)max(end - beg, 1);
                r.x /= cnt; r.y /= cnt; r.z /= cnt; r.w /= cnt;
            }
            float4* o = reinterpret_cast<float4*>(out + (long long)s * ldo) + c;
            if (accumulate) {
                float4 t = *o;
                r.x += t.x; r.y += t.y; r.z += t.z; r.w += t.w;
            }
            *o = r;
        }
        if (amax_out) {          // block-wide max |out[s][:]| (row rescaling of the tensor-core GEMM that consumes it)
            float mx = fmaxf(fmaxf(fabsf(r.x), fabsf(r.y)), fmaxf(fabsf(r.z), fabsf(r.w)));
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            if ((threadIdx.x & 31) == 0) wm[threadIdx.x >> 5] = mx;
            __syncthreads();
            if (threadIdx.x == 0) {
                mx = fmaxf(fmaxf(wm[0], wm[1]), fmaxf(wm[2], wm[3]));
                atomicMax(reinterpret_cast<unsigned*>(amax_out + s), __float_as_uint(mx));
            }
            __syncthreads();
        }
    }
}


__global__ void gather_rows_dsilu_kernel(const float* __restrict__ dOut, int ldd, const int* __restrict__ idx,
                                         const int* __restrict__ ptr, const float* __restrict__ z, int ldz,
                                         float* __restrict__ dX, int ldx, int E, int H4, float* __restrict__ amax_out) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)E * H4) return;
    int e = (int)(t / H4), c = (int)(t - (long long)e * H4);
    int s = idx ? __ldg(idx + e) : e;
    float4 g = __ldg(reinterpret_cast<const float4*>(dOut + (long long)s * ldd) + c);
    if (ptr) {
        float cnt = (float)max(__ldg(ptr + s + 1) - __ldg(ptr + s), 1);
        g.x /= cnt; g.y /= cnt; g.z /= cnt; g.w /= cnt;
    }
    if (z) {
        float4 zz = __ldcs(reinterpret_cast<const float4*>(z + (long long)e * ldz) + c);
        g.x *= mi_dsilu(zz.x); g.y *= mi_dsilu(zz.y); g.z *= mi_dsilu(zz.z); g.w *= mi_dsilu(zz.w);
    }
    *(reinterpret_cast<float4*>(dX + (long long)e * ldx) + c) = g;
    if (amax_out) {          // row maxima for the tensor-core GEMM that consumes dX (gradients span many binades)
        float m = fmaxf(fmaxf(fabsf(g.x), fabsf(g.y)), fmaxf(fabsf(g.z), fabsf(g.w)));
        if ((H4 & 31) == 0) {                     // a warp stays inside one row
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
            if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<unsigned*>(amax_out + e), __float_as_uint(m));
        } else {
            atomicMax(reinterpret_cast<unsigned*>(amax_out + e), __float_as_uint(m));
        }
    }
}

// out[n] += sum over a slab of rows; block (32, 8), grid (ceil(N/32), slabs)
__global__ void colsum_kernel(const float* __restrict__ X, int ldx, int M, int N, int rows_per_block,
                              float* __restrict__ out) {
    __shared__ float red[8][33];
    int col = blockIdx.x * 32 + threadIdx.x;
    int m0 = blockIdx.y * rows_per_block;
    int m1 = min(M, m0 + rows_per_block);
    float acc = 0.f;
    if (col < N)
        for (int m = m0 + threadIdx.y; m < m1; m += 8) acc += __ldg(X + (long long)m * ldx + col);
    red[threadIdx.y][threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.y == 0 && col < N) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) s += red[i][threadIdx.x];
        atomicAdd(out + col, s);
    }
}

// ------------------------------------------------------------------------------------ output heads (inference)
// final LayerNorm + coord_out + type_out + graph mean + lattice_out (+ L product) of cspnet.py:276-294 in ONE launch, one
// CTA per crystal.  As separate launches these were 3 (corrector) / 7 (predictor) latency-bound kernels of ~15 us each for
// 0.3 GFLOP; here the crystal's rows are normalised into shared memory once (same arithmetic as layernorm_fwd_kernel) and
// every warp takes output columns: the weight row sits in registers, the rows come from shared memory.
// NT threads: 256 with two CTAs per SM; 512 (one per SM) when there are no more crystals than SMs — the kernel is a chain of
// L2 round trips per warp (LayerNorm rows, then the weight rows of four output columns at a time), twice the warps halve it.
constexpr int HEAD_ROWS = 32;                       // rows per pass through shared memory
template <int NV, int NT>                           // H = 32 * NV
__global__ void __launch_bounds__(NT, NT == 256 ? 2 : 1) output_heads_kernel(const float* __restrict__ h, int ldh, const int* __restrict__ node_off,
                                                              const float* __restrict__ ln_g, const float* __restrict__ ln_b,
                                                              float eps, const float* __restrict__ coord_w, float* __restrict__ pred_x,
                                                              const float* __restrict__ type_w, const float* __restrict__ type_b, int A,
                                                              float* __restrict__ pred_a, const float* __restrict__ lattice_w,
                                                              const float* __restrict__ L, int ip, float* __restrict__ pred_l) {
    constexpr int H = 32 * NV;
    extern __shared__ float hs[];                    // [HEAD_ROWS][H] normalised rows | [H] column sums | [9]
    float* colsum = hs + HEAD_ROWS * H;
    float* lat9 = colsum + H;
    const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    const int r0 = __ldg(node_off + b), r1 = __ldg(node_off + b + 1);
    for (int c = threadIdx.x; c < H; c += blockDim.x) colsum[c] = 0.f;
    const int n_out = (pred_x ? 3 : 0) + (pred_a ? A : 0);
    for (int c0 = r0; c0 < r1; c0 += HEAD_ROWS) {
        const int nrow = min(HEAD_ROWS, r1 - c0);
        __syncthreads();
        for (int r = warp; r < HEAD_ROWS; r += nw) {  // LayerNorm (or a plain copy) of one row per warp; rows past the crystal: zeros
            float xv[NV];
            if (r < nrow) {
                const float* xr = h + (long long)(c0 + r) * ldh;
                float sm = 0.f;
#pragma unroll
                for (int i = 0; i < NV; ++i) { xv[i] = xr[lane + 32 * i]; sm += xv[i]; }
                if (ln_g) {
                    sm = mi_warp_sum(sm);
                    const float mean = sm / (float)H;
                    float v = 0.f;
#pragma unroll
                    for (int i = 0; i < NV; ++i) { const float d = xv[i] - mean; v += d * d; }
                    v = mi_warp_sum(v);
                    const float rstd = 1.0f / sqrtf(v / (float)H + eps);
#pragma unroll
                    for (int i = 0; i < NV; ++i) xv[i] = (xv[i] - mean) * rstd * __ldg(ln_g + lane + 32 * i) + __ldg(ln_b + lane + 32 * i);
                }
            } else {
#pragma unroll
                for (int i = 0; i < NV; ++i) xv[i] = 0.f;
            }
#pragma unroll
            for (int i = 0; i < NV; ++i) hs[r * H + lane + 32 * i] = xv[i];
        }
        __syncthreads();
        if (pred_l)
            for (int c = threadIdx.x; c < H; c += blockDim.x) {
                float a = colsum[c];
                for (int r = 0; r < nrow; ++r) a += hs[r * H + c];
                colsum[c] = a;
            }
        auto wrow = [&](int o) {                       // weight row of output column o (coordinate head first); past the end: row 0
            const bool is_x_ = pred_x && o < 3;
            const int oc_ = is_x_ ? o : o - (pred_x ? 3 : 0);
            if (o >= n_out) return pred_x ? coord_w : type_w;
            return is_x_ ? coord_w + (long long)oc_ * H : type_w + (long long)oc_ * H;
        };
        const int nrow4 = (nrow + 3) & ~3;
        // register tile: 4 output columns x 4 rows per warp step — every row is read from shared memory once per FOUR
        // output columns (one column at a time made this kernel shared-memory bound: 64 loads per 64 FMAs)
        for (int ob = warp * 4; ob < n_out; ob += nw * 4) {
            float wv[4][NV];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const float* wr = wrow(ob + c);
#pragma unroll
                for (int i = 0; i < NV; ++i) wv[c][i] = __ldg(wr + lane + 32 * i);
            }
            for (int r = 0; r < nrow4; r += 4) {
                float d[16];
#pragma unroll
                for (int k = 0; k < 16; ++k) d[k] = 0.f;
                const float* hr = hs + r * H + lane;
#pragma unroll
                for (int i = 0; i < NV; ++i) {
                    const float x0 = hr[32 * i], x1 = hr[H + 32 * i], x2 = hr[2 * H + 32 * i], x3 = hr[3 * H + 32 * i];
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        d[c] = fmaf(x0, wv[c][i], d[c]);
                        d[4 + c] = fmaf(x1, wv[c][i], d[4 + c]);
                        d[8 + c] = fmaf(x2, wv[c][i], d[8 + c]);
                        d[12 + c] = fmaf(x3, wv[c][i], d[12 + c]);
                    }
                }
                // reduce-scatter of the 16 sums over the lanes: lane l ends with sum number (l >> 1) & 15 = row * 4 + column
                float e8[8], e4[4], e2[2];
                const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4, b1 = lane & 2;
#pragma unroll
                for (int k = 0; k < 8; ++k) e8[k] = (b4 ? d[8 + k] : d[k]) + __shfl_xor_sync(0xffffffffu, b4 ? d[k] : d[8 + k], 16);
#pragma unroll
                for (int k = 0; k < 4; ++k) e4[k] = (b3 ? e8[4 + k] : e8[k]) + __shfl_xor_sync(0xffffffffu, b3 ? e8[k] : e8[4 + k], 8);
#pragma unroll
                for (int k = 0; k < 2; ++k) e2[k] = (b2 ? e4[2 + k] : e4[k]) + __shfl_xor_sync(0xffffffffu, b2 ? e4[k] : e4[2 + k], 4);
                float u = (b1 ? e2[1] : e2[0]) + __shfl_xor_sync(0xffffffffu, b1 ? e2[0] : e2[1], 2);
                u += __shfl_xor_sync(0xffffffffu, u, 1);
                const int idx = (lane >> 1) & 15, rr = r + (idx >> 2), o = ob + (idx & 3);
                if ((lane & 1) == 0 && rr < nrow && o < n_out) {
                    const bool is_x = pred_x && o < 3;
                    const int oc = is_x ? o : o - (pred_x ? 3 : 0);
                    if (is_x) pred_x[(long long)(c0 + rr) * 3 + oc] = u;
                    else pred_a[(long long)(c0 + rr) * A + oc] = u + (type_b ? __ldg(type_b + oc) : 0.f);
                }
            }
        }
    }
    if (pred_l) {
        __syncthreads();
        const float inv = 1.0f / (float)max(r1 - r0, 1);
        for (int o = warp; o < 9; o += nw) {
            float d = 0.f;
            for (int c = lane; c < H; c += 32) d = fmaf(colsum[c] * inv, __ldg(lattice_w + (long long)o * H + c), d);
            d = mi_warp_sum(d);
            if (lane == 0) lat9[o] = d;
        }
        __syncthreads();
        if (threadIdx.x < 9) {
            const int r = threadIdx.x / 3, c = threadIdx.x % 3;
            float v = lat9[threadIdx.x];
            if (ip) {                                 // pred_l = lattice_out(mean) (3x3) @ L[b]   (cspnet.py:288-289)
                const float* l = L + 9 * (long long)b;
                v = 0.f;
#pragma unroll
                for (int k = 0; k < 3; ++k) v += lat9[3 * r + k] * __ldg(l + 3 * k + c);
            }
            pred_l[9 * (long long)b + threadIdx.x] = v;
        }
    }
}

// out[r] = max_c |X[r][c]|  (one warp per row): row maxima of an operand whose producer does not report them
__global__ void row_amax_kernel(const float* __restrict__ X, int ldx, int rows, int cols, float* __restrict__ out) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* xr = X + (long long)row * ldx;
    float mx = 0.f;
    for (int c = lane; c < cols; c += 32) mx = fmaxf(mx, fabsf(xr[c]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) out[row] = mx;
}

// ------------------------------------------------------------------------------------ LayerNorm
// one warp per row
__global__ void layernorm_fwd_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ gamma,
                                     const float* __restrict__ beta, float* __restrict__ y, int ldy,
                                     float* __restrict__ mean_o, float* __restrict__ rstd_o, int rows, int H,
                                     float eps, float* __restrict__ amax_out) {
    int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* xr = x + (long long)row * ldx;
    float s = 0.f;
    for (int c = lane; c < H; c += 32) s += xr[c];
    s = mi_warp_sum(s);
    float mean = s / (float)H;
    float v = 0.f;
    for (int c = lane; c < H; c += 32) { float d = xr[c] - mean; v += d * d; }
    v = mi_warp_sum(v);
    float rstd = 1.0f / sqrtf(v / (float)H + eps);
    float* yr = y + (long long)row * ldy;
    float mx = 0.f;
    for (int c = lane; c < H; c += 32) {
        float yv = (xr[c] - mean) * rstd * __ldg(gamma + c) + __ldg(beta + c);
        yr[c] = yv;
        mx = fmaxf(mx, fabsf(yv));
    }
    if (amax_out) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if (lane == 0) atomicMax(reinterpret_cast<unsigned*>(amax_out + row), __float_as_uint(mx));
    }
    if (lane == 0) {
        if (mean_o) mean_o[row] = mean;
        if (rstd_o) rstd_o[row] = rstd;
    }
}

// LayerNorm whose output feeds a tensor-core GEMM directly: one warp per row, the normalised row is written as the
// pre-split fp16 operand pair of mi_tc_gemm_presplit (two-accumulator format: hi = fp16(s y), lo = fp16((s y - hi) 2^11))
// with s = 2^-e the power of two the GEMM derives from the row maximum (same formula as tc_gemm_kernel: exponent of
// max |y| minus 14, clamped), the row maximum itself (plain store: this kernel owns the row) and, optionally, the
// fp32 row (training keeps it for the backward) and a zeroed companion row (the destination of the fused scatter-mean
// that follows in the layer).  H <= 32 * LN_MAXV.
constexpr int LN_MAXV = 32;
__global__ void layernorm_split_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ gamma,
                                       const float* __restrict__ beta, float* __restrict__ y, int ldy,
                                       __half* __restrict__ y_hi, __half* __restrict__ y_lo, int ldh,
                                       float* __restrict__ amax_o, float* __restrict__ zero_o, int ldz, int zero_cols,
                                       float* __restrict__ mean_o, float* __restrict__ rstd_o, int rows, int H, float eps) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* xr = x + (long long)row * ldx;
    float xv[LN_MAXV];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i) {
        const int c = lane + 32 * i;
        xv[i] = c < H ? xr[c] : 0.f;
        s += xv[i];
    }
    s = mi_warp_sum(s);
    const float mean = s / (float)H;
    float v = 0.f;
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i) {
        const float d = (lane + 32 * i < H) ? xv[i] - mean : 0.f;
        v += d * d;
    }
    v = mi_warp_sum(v);
    const float rstd = 1.0f / sqrtf(v / (float)H + eps);
    float mx = 0.f;
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i) {
        const int c = lane + 32 * i;
        if (c < H) {
            xv[i] = (xv[i] - mean) * rstd * __ldg(gamma + c) + __ldg(beta + c);
            mx = fmaxf(mx, fabsf(xv[i]));
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    const int ex = (int)((__float_as_uint(mx) >> 23) & 0xff) - 127;
    const int e8 = max(-100, min(ex - 14, 100));
    const float sc = __uint_as_float((uint32_t)(127 - e8) << 23);       // 2^-e8, exact
    __half* hr = y_hi + (long long)row * ldh;
    __half* lr = y_lo + (long long)row * ldh;
    float* yr = y ? y + (long long)row * ldy : nullptr;
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i) {
        const int c = lane + 32 * i;
        if (c < H) {
            const float t = xv[i] * sc;
            const __half h = __float2half_rn(t);
            hr[c] = h;
            lr[c] = __float2half_rn((t - __half2float(h)) * 2048.0f);
            if (yr) yr[c] = xv[i];
        }
    }
    if (zero_o) {
        float* zr = zero_o + (long long)row * ldz;
        for (int c = lane; c < zero_cols; c += 32) zr[c] = 0.f;
    }
    if (lane == 0) {
        amax_o[row] = mx;
        if (mean_o) mean_o[row] = mean;
        if (rstd_o) rstd_o[row] = rstd;
    }
}

__global__ void layernorm_bwd_dx_kernel(const float* __restrict__ dy, int lddy, const float* __restrict__ x,
                                        int ldx, const float* __restrict__ gamma, const float* __restrict__ mean,
                                        const float* __restrict__ rstd, float* __restrict__ dx, int lddx,
                                        int accumulate, int rows, int H) {
    int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* xr = x + (long long)row * ldx;
    const float* dyr = dy + (long long)row * lddy;
    float mu = mean[row], rs = rstd[row];
    float c1 = 0.f, c2 = 0.f;
    for (int c = lane; c < H; c += 32) {
        float g = dyr[c] * __ldg(gamma + c);
        float xh = (xr[c] - mu) * rs;
        c1 += g;
        c2 += g * xh;
    }
    c1 = mi_warp_sum(c1) / (float)H;
    c2 = mi_warp_sum(c2) / (float)H;
    float* dxr = dx + (long long)row * lddx;
    for (int c = lane; c < H; c += 32) {
        float g = dyr[c] * __ldg(gamma + c);
        float xh = (xr[c] - mu) * rs;
        float r = rs * (g - c1 - xh * c2);
        dxr[c] = accumulate ? dxr[c] + r : r;
    }
}

// dgamma[c] += sum_rows dy*xhat ; dbeta[c] += sum_rows dy ; block (32, 8)
__global__ void layernorm_bwd_param_kernel(const float* __restrict__ dy, int lddy, const float* __restrict__ x,
                                           int ldx, const float* __restrict__ mean, const float* __restrict__ rstd,
                                           float* __restrict__ dgamma, float* __restrict__ dbeta, int rows, int H,
                                           int rows_per_block) {
    __shared__ float rg[8][33], rb[8][33];
    int col = blockIdx.x * 32 + threadIdx.x;
    int m0 = blockIdx.y * rows_per_block, m1 = min(rows, m0 + rows_per_block);
    float ag = 0.f, ab = 0.f;
    if (col < H)
        for (int m = m0 + threadIdx.y; m < m1; m += 8) {
            float d = __ldg(dy + (long long)m * lddy + col);
            float xh = (__ldg(x + (long long)m * ldx + col) - __ldg(mean + m)) * __ldg(rstd + m);
            ag += d * xh;
            ab += d;
        }
    rg[threadIdx.y][threadIdx.x] = ag;
    rb[threadIdx.y][threadIdx.x] = ab;
    __syncthreads();
    if (threadIdx.y == 0 && col < H) {
        float sg = 0.f, sb = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) { sg += rg[i][threadIdx.x]; sb += rb[i][threadIdx.x]; }
        atomicAdd(dgamma + col, sg);
        atomicAdd(dbeta + col, sb);
    }
}

// ------------------------------------------------------------------------------------ per-crystal
__global__ void lattice_ip_kernel(const float* __restrict__ L, float* __restrict__ ips, int B) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B * 9) return;
    int b = t / 9, r = (t % 9) / 3, c = t % 3;
    const float* l = L + 9 * b;
    ips[t] = l[3 * r] * l[3 * c] + l[3 * r + 1] * l[3 * c + 1] + l[3 * r + 2] * l[3 * c + 2];
}
// out[b][n] = bias[n] + sum_k vec(L_b L_b^T)[k] * W[n][k]   (the lattice part of the first edge linear, cspnet.py:67-72)
// blockIdx.y = weight set (layer): W, bias and out advance by their strides, the lattices are shared
__global__ void lattice_linear_kernel(const float* __restrict__ L, const float* __restrict__ W, const float* __restrict__ bias,
                                      float* __restrict__ out, int ldo, int H, long long w_stride, long long b_stride,
                                      long long out_stride) {
    __shared__ float ip[9];
    const int b = blockIdx.x;
    W += blockIdx.y * w_stride;
    if (bias) bias += blockIdx.y * b_stride;
    out += blockIdx.y * out_stride;
    if (threadIdx.x < 9) {
        const float* l = L + 9 * b;
        int r = threadIdx.x / 3, c = threadIdx.x % 3;
        ip[threadIdx.x] = l[3 * r] * l[3 * c] + l[3 * r + 1] * l[3 * c + 1] + l[3 * r + 2] * l[3 * c + 2];
    }
    __syncthreads();
    for (int n = threadIdx.x; n < H; n += blockDim.x) {
        const float* w = W + 9 * n;
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < 9; ++k) acc = fmaf(ip[k], __ldg(w + k), acc);
        out[(long long)b * ldo + n] = acc + (bias ? __ldg(bias + n) : 0.f);
    }
}
__global__ void bmm3_kernel(const float* __restrict__ A, const float* __restrict__ L, float* __restrict__ out,
                            int B, int transL) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B * 9) return;
    int b = t / 9, r = (t % 9) / 3, c = t % 3;
    const float* a = A + 9 * b;
    const float* l = L + 9 * b;
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k) s += a[3 * r + k] * (transL ? l[3 * c + k] : l[3 * k + c]);
    out[t] = s;
}
__global__ void time_embed_kernel(const int* __restrict__ t, const float* __restrict__ freq, int B, int half,
                                  float* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * half) return;
    int b = i / half, k = i - b * half;
    float arg = __fmul_rn((float)__ldg(t + b), __ldg(freq + k));
    float s, c;
    sincosf(arg, &s, &c);
    out[(long long)b * 2 * half + k] = s;
    out[(long long)b * 2 * half + half + k] = c;
}
__global__ void lattice_params_to_matrix_kernel(const float* __restrict__ lengths, const float* __restrict__ angles,
                                                float* __restrict__ L, int B) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const float d2r = 0.017453292519943295f;
    float ca[3], sa[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        float r = angles[3 * b + i] * d2r;
        ca[i] = cosf(r);
        sa[i] = sinf(r);
    }
    float val = (ca[0] * ca[1] - ca[2]) / (sa[0] * sa[1]);
    val = fminf(fmaxf(val, -1.f), 1.f);
    float gs = acosf(val);
    float l0 = lengths[3 * b], l1 = lengths[3 * b + 1], l2 = lengths[3 * b + 2];
    float* o = L + 9 * b;
    o[0] = l0 * sa[1]; o[1] = 0.f; o[2] = l0 * ca[1];
    o[3] = -l1 * sa[0] * cosf(gs); o[4] = l1 * sa[0] * sinf(gs); o[5] = l1 * ca[0];
    o[6] = 0.f; o[7] = 0.f; o[8] = l2;
}
__global__ void lattice_matrix_to_params_kernel(const float* __restrict__ L, float* __restrict__ lengths,
                                                float* __restrict__ angles, int B) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const float* l = L + 9 * b;
    float len[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) len[i] = sqrtf(l[3 * i] * l[3 * i] + l[3 * i + 1] * l[3 * i + 1] + l[3 * i + 2] * l[3 * i + 2]);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        int j = (i + 1) % 3, k = (i + 2) % 3;
        float dot = l[3 * j] * l[3 * k] + l[3 * j + 1] * l[3 * k + 1] + l[3 * j + 2] * l[3 * k + 2];
        float c = fminf(fmaxf(dot / (len[j] * len[k]), -1.f), 1.f);
        angles[3 * b + i] = acosf(c) * 180.0f / 3.14159265358979323846f;
        lengths[3 * b + i] = len[i];
    }
}
__global__ void argmax_rows_kernel(const float* __restrict__ a, int lda, int rows, int cols, int add,
                                   int* __restrict__ out) {
    int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* r = a + (long long)row * lda;
    float best = -INFINITY;
    int bi = 0x7fffffff;
    for (int c = lane; c < cols; c += 32) {
        float v = r[c];
        if (v > best || (v == best && c < bi) || (v != v && best == best)) { best = v; bi = c; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        float ov = __shfl_xor_sync(0xffffffffu, best, o);
        int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if (lane == 0) out[row] = bi + add;
}

// ------------------------------------------------------------------------------------ reverse updates
// coef row layout (8 floats per t): sqrt_sn, step_c, std_c, step_p, std_p, c0, c1, sig
__global__ void reverse_corrector_kernel(const float* __restrict__ x, const float* __restrict__ pred_x,
                                         const float* __restrict__ z, float* __restrict__ x_half, int n,
                                         const float* __restrict__ coef, const int* __restrict__ t_dev, int t_host) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* c = coef + 8 * (t_dev ? __ldg(t_dev) : t_host);
    const float sqrt_sn = __ldg(c + 0), step_c = __ldg(c + 1), std_c = __ldg(c + 2);
    float px = __fmul_rn(pred_x[i], sqrt_sn);
    float v = __fsub_rn(x[i], __fmul_rn(step_c, px));
    float zz = z ? z[i] : 0.f;
    x_half[i] = __fadd_rn(v, __fmul_rn(std_c, zz));
}
__global__ void reverse_predictor_kernel(const float* __restrict__ x_half, const float* __restrict__ pred_x,
                                         const float* __restrict__ z_x, float* __restrict__ x, int n3,
                                         float* __restrict__ l, const float* __restrict__ pred_l,
                                         const float* __restrict__ z_l, int n9, float* __restrict__ a,
                                         const float* __restrict__ pred_a, const float* __restrict__ z_a,
                                         long long na, const float* __restrict__ coef,
                                         const int* __restrict__ t_dev, int t_host) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const float* c = coef + 8 * (t_dev ? __ldg(t_dev) : t_host);
    const float sqrt_sn = __ldg(c + 0), step_p = __ldg(c + 3), std_p = __ldg(c + 4);
    const float c0 = __ldg(c + 5), c1 = __ldg(c + 6), sig = __ldg(c + 7);
    if (i < na) {
        float v = __fmul_rn(c0, __fsub_rn(a[i], __fmul_rn(c1, pred_a[i])));
        a[i] = __fadd_rn(v, __fmul_rn(sig, z_a ? z_a[i] : 0.f));
    }
    if (i < n3) {
        float px = __fmul_rn(pred_x[i], sqrt_sn);
        float v = __fsub_rn(x_half[i], __fmul_rn(step_p, px));
        v = __fadd_rn(v, __fmul_rn(std_p, z_x ? z_x[i] : 0.f));
        x[i] = mi_mod1(mi_mod1(v));
    }
    if (i < n9) {
        float v = __fmul_rn(c0, __fsub_rn(l[i], __fmul_rn(c1, pred_l[i])));
        l[i] = __fadd_rn(v, __fmul_rn(sig, z_l ? z_l[i] : 0.f));
    }
}
__global__ void step_begin_kernel(const int* __restrict__ t_dev, const float* __restrict__ ttab,
                                  float* __restrict__ temb, int B, int T) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * T) return;
    int k = i % T;
    temb[i] = __ldg(ttab + (long long)__ldg(t_dev) * T + k);
}
__global__ void step_end_kernel(int* t_dev) { *t_dev -= 1; }

// ------------------------------------------------------------------------------------ add_noise
__device__ __forceinline__ float wn_score_ref(float x, float sigma) {
    // scheduler.py:32-43, N = 10, T = 1, same operation order in fp32
    const float s2 = __fmul_rn(sigma, sigma);
    float p = 0.f, q = 0.f;
    for (int i = -10; i <= 10; ++i) {
        float t = __fadd_rn(x, (float)i);
        float ex = expf(__fdiv_rn(__fdiv_rn(-__fmul_rn(t, t), 2.f), s2));
        p = __fadd_rn(p, ex);
        q = __fadd_rn(q, __fmul_rn(__fdiv_rn(t, s2), ex));
    }
    return __fdiv_rn(q, p);
}
__global__ void add_noise_kernel(const float* __restrict__ L0, const float* __restrict__ x0,
                                 const int* __restrict__ Z, const float* __restrict__ z_l,
                                 const float* __restrict__ z_x, const float* __restrict__ z_a, int B, int N,
                                 int A, const float* __restrict__ coef, const int* __restrict__ t_dev, int t_host,
                                 float* __restrict__ l_t, float* __restrict__ x_t, float* __restrict__ a_t,
                                 float* __restrict__ tar_x) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const float* cf = coef + 4 * (t_dev ? __ldg(t_dev) : t_host);     // {c0, c1, sigma, sqrt(sigma_norm)}
    const float c0 = __ldg(cf), c1 = __ldg(cf + 1), sigma = __ldg(cf + 2), sqrt_sn = __ldg(cf + 3);
    if (i < (long long)N * A) {
        int node = (int)(i / A), k = (int)(i - (long long)node * A);
        float oh = (k == __ldg(Z + node) - 1) ? 1.f : 0.f;
        a_t[i] = __fadd_rn(__fmul_rn(c0, oh), __fmul_rn(c1, z_a[i]));
    }
    if (i < 3LL * N) {
        float nz = __fmul_rn(sigma, z_x[i]);
        x_t[i] = mi_mod1(__fadd_rn(x0[i], nz));
        tar_x[i] = __fdiv_rn(wn_score_ref(nz, sigma), sqrt_sn);
    }
    if (i < 9LL * B) l_t[i] = __fadd_rn(__fmul_rn(c0, L0[i]), __fmul_rn(c1, z_l[i]));
}

// ------------------------------------------------------------------------------------ losses
// one block (128 threads) per crystal
__global__ void __launch_bounds__(128) rl_loss_kernel(
    const float* __restrict__ pred_l, const float* __restrict__ pred_x, const float* __restrict__ pred_a,
    const float* __restrict__ tgt_l, const float* __restrict__ tgt_x, const float* __restrict__ tgt_a,
    const float* __restrict__ prior_l, const float* __restrict__ prior_x, const float* __restrict__ prior_a,
    const int* __restrict__ node_off, int A, float cost_l, float cost_x, float cost_a,
    const float* __restrict__ w_loss, const float* __restrict__ w_kl, float scale, float* __restrict__ loss,
    float* __restrict__ kl, float* __restrict__ d_l, float* __restrict__ d_x, float* __restrict__ d_a,
    float* __restrict__ stats) {
    const int b = blockIdx.x, tid = threadIdx.x;
    const int n0 = node_off[b], n = node_off[b + 1] - n0;
    const float wl = w_loss ? w_loss[b] * scale : 0.f;
    const float wk = (w_kl && prior_l) ? w_kl[b] * scale : 0.f;
    const float inv_n = 1.f / (float)max(n, 1);
    // s[0..2]: sum sq (l, x, a) vs target ; s[3..5]: vs prior
    float s[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int i = tid; i < 9; i += 128) {
        float p = pred_l[9 * b + i];
        float dt = tgt_l ? p - tgt_l[9 * b + i] : 0.f;
        float dp = prior_l ? p - prior_l[9 * b + i] : 0.f;
        s[0] += dt * dt;
        s[3] += dp * dp;
        if (d_l) d_l[9 * b + i] = wl * cost_l * 2.f * dt / 9.f + wk * 2.f * dp / 9.f;
    }
    for (int i = tid; i < 3 * n; i += 128) {
        long long g = 3LL * n0 + i;
        float p = pred_x[g];
        float dt = tgt_x ? p - tgt_x[g] : 0.f;
        float dp = prior_x ? p - prior_x[g] : 0.f;
        s[1] += dt * dt;
        s[4] += dp * dp;
        if (d_x) d_x[g] = (wl * cost_x * 2.f * dt + wk * 2.f * dp) * inv_n / 3.f;
    }
    const float inv_A = 1.f / (float)A;
    for (int i = tid; i < A * n; i += 128) {
        long long g = (long long)A * n0 + i;
        float p = pred_a[g];
        float dt = tgt_a ? p - tgt_a[g] : 0.f;
        float dp = prior_a ? p - prior_a[g] : 0.f;
        s[2] += dt * dt;
        s[5] += dp * dp;
        if (d_a) d_a[g] = (wl * cost_a * 2.f * dt + wk * 2.f * dp) * inv_n * inv_A;
    }
    __shared__ float red[4][6];
#pragma unroll
    for (int q = 0; q < 6; ++q) s[q] = mi_warp_sum(s[q]);
    if ((tid & 31) == 0)
#pragma unroll
        for (int q = 0; q < 6; ++q) red[tid >> 5][q] = s[q];
    __syncthreads();
    if (tid == 0) {
        float t[6];
#pragma unroll
        for (int q = 0; q < 6; ++q) t[q] = red[0][q] + red[1][q] + red[2][q] + red[3][q];
        if (loss) loss[b] = cost_l * (t[0] / 9.f) + cost_x * (t[1] / 3.f * inv_n) + cost_a * (t[2] * inv_A * inv_n);
        float klb = t[3] / 9.f + t[4] / 3.f * inv_n + t[5] * inv_A * inv_n;
        if (kl) kl[b] = klb;
        if (stats) {   // running sums for the ft_step logs (pipeline/mat_invent.py:168-170)
            float lb = cost_l * (t[0] / 9.f) + cost_x * (t[1] / 3.f * inv_n) + cost_a * (t[2] * inv_A * inv_n);
            atomicAdd(stats + 0, (w_loss ? w_loss[b] : 0.f) * lb);
            atomicAdd(stats + 1, ((w_kl && prior_l) ? w_kl[b] : 0.f) * klb);
        }
    }
}

// ------------------------------------------------------------------------------------ Adam
__global__ void adam_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, long long n, float step_size, float omb1, float b2, float omb2,
                            float eps, float bc2_sqrt, float grad_scale, int zero_grad) {
    long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i >= n) return;
    if (i + 4 <= n) {
        float4 P = *reinterpret_cast<float4*>(p + i), G = *reinterpret_cast<float4*>(g + i);
        float4 Mm = *reinterpret_cast<float4*>(m + i), V = *reinterpret_cast<float4*>(v + i);
        float pp[4] = {P.x, P.y, P.z, P.w}, gg[4] = {G.x, G.y, G.z, G.w};
        float mm[4] = {Mm.x, Mm.y, Mm.z, Mm.w}, vv[4] = {V.x, V.y, V.z, V.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float gr = gg[j] * grad_scale;
            mm[j] = mm[j] + (gr - mm[j]) * omb1;          // exp_avg.lerp_(grad, 1 - beta1)
            vv[j] = b2 * vv[j] + omb2 * gr * gr;
            float denom = sqrtf(vv[j]) / bc2_sqrt + eps;
            pp[j] = pp[j] - step_size * (mm[j] / denom);
        }
        *reinterpret_cast<float4*>(p + i) = make_float4(pp[0], pp[1], pp[2], pp[3]);
        *reinterpret_cast<float4*>(m + i) = make_float4(mm[0], mm[1], mm[2], mm[3]);
        *reinterpret_cast<float4*>(v + i) = make_float4(vv[0], vv[1], vv[2], vv[3]);
        if (zero_grad) *reinterpret_cast<float4*>(g + i) = make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
        for (long long j = i; j < n; ++j) {
            float gr = g[j] * grad_scale;
            float mj = m[j] + (gr - m[j]) * omb1;
            float vj = b2 * v[j] + omb2 * gr * gr;
            float denom = sqrtf(vj) / bc2_sqrt + eps;
            p[j] = p[j] - step_size * (mj / denom);
            m[j] = mj;
            v[j] = vj;
            if (zero_grad) g[j] = 0.f;
        }
    }
}

// ------------------------------------------------------------------------------------ Philox4x32-10
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
    const unsigned M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        unsigned hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
        unsigned hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += W0;
        key.y += W1;
    }
    return ctr;
}
__device__ __forceinline__ float u01(unsigned x) { return ((float)(x >> 8) + 0.5f) * (1.0f / 16777216.0f); }

template <bool NORMAL>
__global__ void philox_fill_kernel(float* __restrict__ out, long long n, unsigned long long seed,
                                   unsigned long long offset, const unsigned long long* __restrict__ offset_dev) {
    long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // group of 4 outputs
    if (q * 4 >= n) return;
    unsigned long long c = offset + (offset_dev ? *offset_dev : 0ull) + (unsigned long long)q;
    uint4 r = philox4x32_10(make_uint4((unsigned)c, (unsigned)(c >> 32), 0u, 0u),
                            make_uint2((unsigned)seed, (unsigned)(seed >> 32)));
    float o[4];
    if (NORMAL) {
        float u0 = u01(r.x), u1 = u01(r.y), u2 = u01(r.z), u3 = u01(r.w);
        float ra = sqrtf(-2.f * logf(u0)), rb = sqrtf(-2.f * logf(u2));
        float s0, c0, s1, c1;
        sincospif(2.f * u1, &s0, &c0);
        sincospif(2.f * u3, &s1, &c1);
        o[0] = ra * c0; o[1] = ra * s0; o[2] = rb * c1; o[3] = rb * s1;
    } else {
        o[0] = (float)(r.x >> 8) * (1.0f / 16777216.0f);
        o[1] = (float)(r.y >> 8) * (1.0f / 16777216.0f);
        o[2] = (float)(r.z >> 8) * (1.0f / 16777216.0f);
        o[3] = (float)(r.w >> 8) * (1.0f / 16777216.0f);
    }
    long long i = q * 4;
    if (i + 4 <= n && mi_aligned16(out + i)) *reinterpret_cast<float4*>(out + i) = make_float4(o[0], o[1], o[2], o[3]);
    else
        for (int j = 0; j < 4 && i + j < n; ++j) out[i + j] = o[j];
}
__global__ void advance_offset_kernel(unsigned long long* off, unsigned long long inc) { *off += inc; }

}  // namespace

// ==================================================================================== C ABI
extern "C" int mi_fc_edges(const int* node_off, const int* edge_off, int B, int N, int E, int* edge_src,
                           int* edge_dst, int* edge_graph, int* seg_ptr, int* dst_ptr, int* dst_perm,
                           int* node_graph, mi_stream_t stream) {
    MI_CHECK_ARG(B >= 0 && N >= 0 && E >= 0, "negative size");
    MI_CHECK_ARG(node_off && edge_off && edge_src && edge_dst && edge_graph && seg_ptr && node_graph, "null pointer");
    if (B == 0) return MI_OK;
    int n = max(max(E, N), 1);
    fc_edges_kernel<<<mi_div_up(n, 256), 256, 0, (cudaStream_t)stream>>>(node_off, edge_off, B, N, E, edge_src, edge_dst,
                                                                         edge_graph, seg_ptr, dst_ptr, dst_perm, node_graph);
    MI_CHECK_LAUNCH();
    return MI_OK;
}

extern "C" int mi_edge_fourier(const float* x, const int* edge_src, const int* edge_dst, const float* cell_off,
                               int E, int F, float* frac_diff, float* phi, int ld_phi, void* phi_hi, void* phi_lo,
                               float op_scale, float lo_scale, mi_stream_t stream) {
    MI_CHECK_ARG(E >= 0 && F > 0 && ld_phi >= 6 * F, "bad sizes");
    if (E == 0) return MI_OK;
    MI_CHECK_ARG(x && edge_src && edge_dst && (phi || phi_hi) && ((phi_hi == nullptr) == (phi_lo == nullptr)), "null pointer");
    const bool vec = (F % 4 == 0) && (ld_phi % 4 == 0) && (!phi || mi_host_aligned16(phi)) && (!phi_hi || (mi_host_aligned16(phi_hi) && mi_host_aligned16(phi_lo)));
    if (vec) {
        long long n = (long long)E * (3 * F / 4);
        edge_fourier_kernel<<<mi_div_up(n, 256), 256, 0, (cudaStream_t)stream>>>(x, edge_src, edge_dst, cell_off, E, F, frac_diff, phi,
                                                                                  ld_phi, (__half*)phi_hi, (__half*)phi_lo, op_scale, lo_scale);
    } else {
        long long n = (long long)E * 3 * F;
        edge_fourier_kernel_scalar<<<mi_div_up(n, 256), 256, 0, (cudaStream_t)stream>>>(x, edge_src, edge_dst, cell_off, E, F, frac_diff,
                                                                                         phi, ld_phi, (__half*)phi_hi, (__half*)phi_lo, op_scale, lo_scale);
    }
    MI_CHECK_LAUNCH();
    return MI_OK;
}

extern "C" int mi_segment_reduce(const float* X, int ldx, const int* ptr, const int* perm, float* out, int ldo,
                                 int S, int H, int mean, int accumulate, float* amax_out, mi_stream_t stream) {
    MI_CHECK_ARG(S >= 0 && H > 0 && H % 4 == 0 && ldx % 4 == 0 && ldo % 4 == 0, "H, ldx, ldo must be multiples of 4");
    if (S == 0) return MI_OK;
    MI_CHECK_ARG(X && ptr && out && mi_host_aligned16(X) && mi_host_aligned16(out), "null or unaligned pointer");
    int H4 = H / 4;
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        MI_CUDA(cudaGetDevice(&dev));
        MI_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    }
    const int gy = mi_div_up(H4, 128);
    int gx = sms * 16 / gy;                       // 16 resident 128-thread blocks per SM: one full wave, grid-stride
    if (gx > S) gx = S;
    if (gx < 1) gx = 1;
    dim3 grid(gx, gy);
    segment_reduce_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(X, ldx, ptr, perm, out, ldo, H4, S, mean, accumulate, amax_out);
    MI_CHECK_LAUNCH();
    return MI_OK;
}

extern "C" int mi_gather_rows_dsilu(const float* dOut, int ldd, const int* idx, const int* ptr, const float* z,
                                    int ldz, float* dX, int ldx, int E, int H, float* amax_out, mi_stream_t stream) {
    MI_CHECK_ARG(E >= 0 && H > 0 && H % 4 == 0 && ldd % 4 == 0 && ldx % 4 == 0 && (!z || ldz % 4 == 0), "sizes must be multiples of 4");
    if (E == 0) return MI_OK;
    MI_CHECK_ARG(dOut && dX, "null pointer");
    long long n = (long long)E * (H / 4);
    gather_rows_dsilu_kernel<<<mi_div_up(n, 256), 256, 0, (cudaStream_t)stream>>>(dOut, ldd, idx, ptr, z, ldz, dX, ldx, E, H / 4, amax_out);
    MI_CHECK_LAUNCH();
    return MI_OK;
}

extern "C" int mi_colsum(const float* X, int ldx, int M, int N, float* out, int accumulate, mi_stream_t stream) {
    MI_CHECK_ARG(M >= 0 && N > 0 && out, "bad arguments");
    cudaStream_t s = (cudaStream_t)stream;
    if (!accumulate) MI_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * N, s));
    if (M == 0) return MI_OK;
    MI_CHECK_ARG(X != nullptr, "null X");
    int rpb = 256;
    dim3 grid(mi_div_up(N, 32), mi_div_up(M, rpb));
    colsum_kernel<<<grid, dim3(32, 8), 0, s>>>(X, ldx, M, N, rpb, out);
    MI_CHECK_LAUNCH();
    return MI_OK;
}

extern "C" int mi_layernorm_fwd(const float* x, int ldx, const float* gamma, const float* beta, float* y, int ldy,
                                float* mean, float* rstd, int rows, int H, float eps, float* amax_out, mi_stream_t stream) {
    MI_CHECK_ARG(rows >= 0 && H > 0, "bad sizes");
    if (rows == 0) return MI_OK;
    MI_CHECK_ARG(x && gamma && beta && y, "null pointer");
    layernorm_fwd_kernel<<<mi_div_up(rows, 8), 256, 0, (cudaStream_t)stream>>>(x, ldx, gamma, beta, y, ldy, mean, rstd, rows, H, eps, amax_out);
    MI_CHECK_LAUNCH();
    return MI_OK;
}

extern "C" int mi_layernorm_fwd_split(const float* x, int ldx, const float* gamma, const float* beta, float* y, int ldy,
                                      void* y_hi, void* y_lo, int ldh, float* amax, float* zero_out, int ldz, int zero_cols,
                                      float* mean, float* rstd, int rows, int H, float eps, mi_stream_t stream) {
    MI_CHECK_ARG(rows >= 0 && H > 0 && H <= 32 * LN_MAXV, "bad sizes (H <= 1024)");
    if (rows == 0) return MI_OK;
    MI_CHECK_ARG(x && gamma && beta && y_hi && y_lo && amax, "null pointer");
    MI_CHECK_ARG(ldh >= H && (!y || ldy >= H) && (!zero_out || ldz >= zero_cols), "leading dimension too small");
    layernorm_split_kernel<<<mi_div_up(rows, 8), 256, 0, (cudaStream_t)stream>>>(
        x, ldx, gamma, beta, y, ldy, (__half*)y_hi, (__half*)y_lo, ldh, amax, zero_out, ldz, zero_cols, mean, rstd, rows, H, eps);
    MI_CHECK_LAUNCH();
    return MI_OK;
}

extern "C" int mi_layernorm_bwd(const float* dy, int lddy, const float* x, int ldx, const float* gamma,
                                const float* mean, const float* rstd, float* dx, int lddx, int accumulate_dx,
                                float* dgamma, float* dbeta, int rows, int H, mi_stream_t stream) {
    MI_CHECK_ARG(rows >= 0 && H > 0, "bad sizes");
    if (rows == 0) return MI_OK;
    MI_CHECK_ARG(dy && x && gamma && mean && rstd, "null pointer");
    cudaStream_t s = (cudaStream_t)stream;
    if (dgamma && dbeta) {
        int rpb = 128;
        dim3 grid(mi_div_up(H, 32), mi_div_up(rows, rpb));
        layernorm_bwd_param_kernel<<<grid, dim3(32, 8), 0, s>>>(dy, lddy, x, ldx, mean, rstd, dgamma, dbeta, rows, H, rpb);
        MI_CHECK_LAUNCH();
    }
    if (dx) {
        layernorm_bwd_dx_kernel<<<mi_div_up(rows, 8), 256, 0, s>>>(dy, lddy, x, ldx, gamma, mean, rstd, dx, lddx, accumulate_dx, rows, H);
        MI_CHECK_LAUNCH();
    }
    return MI_OK;
}

extern "C" int mi_lattice_ip(const float* L, float* ips, int B, mi_stream_t stream) {
    if (B <= 0) return MI_OK;
    MI_CHECK_ARG(L && ips, "null pointer");
    lattice_ip_kernel<<<mi_div_up(9 * B, 256), 256, 0, (cudaStream_t)stream>>>(L, ips, B);
    MI_CHECK_LAUNCH();
    return MI_OK;
}
extern "C" int mi_lattice_linear(const float* L, const float* W, const float* bias, float* out, int ldo, int B, int H,
                                 int n_sets, long long w_stride, long long bias_stride, long long out_stride,
                                 mi_stream_t stream) {
    if (B <= 0 || n_sets <= 0) return MI_OK;
    MI_CHECK_ARG(L && W && out && H > 0 && ldo >= H, "bad arguments");
    lattice_linear_kernel<<<dim3(B, n_sets), 128, 0, (cudaStream_t)stream>>>(L, W, bias, out, ldo, H, w_stride, bias_stride, out_stride);
    MI_CHECK_LAUNCH();
    return MI_OK;
}
template <int NV>
static int launch_heads(const float* h, int ldh, const int* node_off, int B, const float* ln_g, const float* ln_b, float eps,
                        const float* coord_w, float* pred_x, const float* type_w, const float* type_b, int A, float* pred_a,
                        const float* lattice_w, const float* L, int ip, float* pred_l, cudaStream_t s) {
    constexpr int H = 32 * NV;
    const size_t smem = ((size_t)HEAD_ROWS * H + H + 16) * sizeof(float);
    static int sms = 0;
    if (sms == 0) {
        MI_CUDA(cudaFuncSetAttribute(output_heads_kernel<NV, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        MI_CUDA(cudaFuncSetAttribute(output_heads_kernel<NV, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int dev = 0;
        MI_CUDA(cudaGetDevice(&dev));
        MI_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    }
    if (B <= sms)
        output_heads_kernel<NV, 512><<<B, 512, smem, s>>>(h, ldh, node_off, ln_g, ln_b, eps, coord_w, pred_x, type_w, type_b, A, pred_a,
                                                          lattice_w, L, ip, pred_l);
    else
        output_heads_kernel<NV, 256><<<B, 256, smem, s>>>(h, ldh, node_off, ln_g, ln_b, eps, coord_w, pred_x, type_w, type_b, A, pred_a,
                                                          lattice_w, L, ip, pred_l);
    MI_CHECK_LAUNCH();
    return MI_OK;
}

extern "C" int mi_output_heads(const float* h, int ldh, const int* node_off, int B, int H, const float* ln_g, const float* ln_b,
                               float eps, const float* coord_w, float* pred_x, const float* type_w, const float* type_b, int A,
                               float* pred_a, const float* lattice_w, const float* L, int ip, float* pred_l, mi_stream_t stream) {
    if (B <= 0) return MI_OK;
    MI_CHECK_ARG(h && node_off && (H == 128 || H == 256 || H == 512 || H == 1024) && ldh >= H, "bad arguments (H in {128, 256, 512, 1024})");
    MI_CHECK_ARG((ln_g == nullptr) == (ln_b == nullptr), "ln_g and ln_b go together");
    MI_CHECK_ARG((!pred_x || coord_w) && (!pred_a || (type_w && A > 0)) && (!pred_l || (lattice_w && (!ip || L))), "null weight");
    cudaStream_t s = (cudaStream_t)stream;
#define MI_HEADS(NV) launch_heads<NV>(h, ldh, node_off, B, ln_g, ln_b, eps, coord_w, pred_x, type_w, type_b, A, pred_a, lattice_w, L, ip, pred_l, s)
    switch (H) {
        case 128: return MI_HEADS(4);
        case 256: return MI_HEADS(8);
        case 512: return MI_HEADS(16);
        default: return MI_HEADS(32);
    }
#undef MI_HEADS
}

extern "C" int mi_row_amax(const float* X, int ldx, int rows, int cols, float* out, mi_stream_t stream) {
    if (rows <= 0) return MI_OK;
    MI_CHECK_ARG(X && out && cols > 0 && ldx >= cols, "bad arguments");
    row_amax_kernel<<<mi_div_up(rows, 8), 256, 0, (cudaStream_t)stream>>>(X, ldx, rows, cols, out);
    MI_CHECK_LAUNCH();
    return MI_OK;
}
extern "C" int mi_bmm3(const float* A, const float* L, float* out, int B, int transL, mi_stream_t stream) {
    if (B <= 0) return MI_OK;
    MI_CHECK_ARG(A && L && out, "null pointer");
    bmm3_kernel<<<mi_div_up(9 * B, 256), 256, 0, (cudaStream_t)stream>>>(A, L, out, B, transL);
    MI_CHECK_LAUNCH();
    return MI_OK;
}
extern "C" int mi_time_embed(const int* t, const float* freq, int B, int dim, float* out, mi_stream_t stream) {
    if (B <= 0) return MI_OK;
    MI_CHECK_ARG(t && freq && out && dim > 0 && dim % 2 == 0, "bad arguments");
    time_embed_kernel<<<mi_div_up((long long)B * dim / 2, 256), 256, 0, (cudaStream_t)stream>>>(t, freq, B, dim / 2, out);
    MI_CHECK_LAUNCH();
    return MI_OK;
}
extern "C" int mi_lattice_params_to_matrix(const float* lengths, const float* angles, float* L, int B, mi_stream_t stream) {
    if (B <= 0) return MI_OK;
    MI_CHECK_ARG(lengths && angles && L, "null pointer");
    lattice_params_to_matrix_kernel<<<mi_div_up(B, 128), 128, 0, (cudaStream_t)stream>>>(lengths, angles, L, B);
    MI_CHECK_LAUNCH();
    return MI_OK;
}
extern "C" int mi_lattice_matrix_to_params(const float* L, float* lengths, float* angles, int B, mi_stream_t stream) {
    if (B <= 0) return MI_OK;
    MI_CHECK_ARG(lengths && angles && L, "null pointer");
    lattice_matrix_to_params_kernel<<<mi_div_up(B, 128), 128, 0, (cudaStream_t)stream>>>(L, lengths, angles, B);
    MI_CHECK_LAUNCH();
    return MI_OK;
}
extern "C" int mi_argmax_rows(const float* a, int lda, int rows, int cols, int add, int* out, mi_stream_t stream) {
    if (rows <= 0) return MI_OK;
    MI_CHECK_ARG(a && out && cols > 0, "bad arguments");
    argmax_rows_kernel<<<mi_div_up(rows, 8), 256, 0, (cudaStream_t)stream>>>(a, lda, rows, cols, add, out);
    MI_CHECK_LAUNCH();
    return MI_OK;
}

extern "C" int mi_reverse_corrector(const float* x, const float* pred_x, const float* z_x, float* x_half, int N,
                                    const float* coef, const int* t_dev, int t_host, mi_stream_t stream) {
    if (N <= 0) return MI_OK;
    MI_CHECK_ARG(x && pred_x && x_half && coef, "null pointer");
    reverse_corrector_kernel<<<mi_div_up(3 * N, 256), 256, 0, (cudaStream_t)stream>>>(x, pred_x, z_x, x_half, 3 * N, coef, t_dev, t_host);
    MI_CHECK_LAUNCH();
    return MI_OK;
}
extern "C" int mi_reverse_predictor(const float* x_half, const float* pred_x, const float* z_x, float* x, int N,
                                    float* l, const float* pred_l, const float* z_l, int B, float* a,
                                    const float* pred_a, const float* z_a, int A, const float* coef,
                                    const int* t_dev, int t_host, mi_stream_t stream) {
    if (N <= 0 || B <= 0) return MI_OK;
    MI_CHECK_ARG(x_half && pred_x && x && l && pred_l && a && pred_a && A > 0 && coef, "null pointer");
    long long na = (long long)N * A;
    long long n = na > 9LL * B ? na : 9LL * B;
    reverse_predictor_kernel<<<mi_div_up(n, 256), 256, 0, (cudaStream_t)stream>>>(
        x_half, pred_x, z_x, x, 3 * N, l, pred_l, z_l, 9 * B, a, pred_a, z_a, na, coef, t_dev, t_host);
    MI_CHECK_LAUNCH();
    return MI_OK;
}
extern "C" int mi_sampler_step_begin(const int* t_dev, const float* ttab, float* temb, int B, int T, mi_stream_t stream) {
    if (B <= 0) return MI_OK;
    MI_CHECK_ARG(t_dev && ttab && temb && T > 0, "null pointer");
    step_begin_kernel<<<mi_div_up((long long)B * T, 256), 256, 0, (cudaStream_t)stream>>>(t_dev, ttab, temb, B, T);
    MI_CHECK_LAUNCH();
    return MI_OK;
}
extern "C" int mi_sampler_step_end(int* t_dev, mi_stream_t stream) {
    MI_CHECK_ARG(t_dev != nullptr, "null pointer");
    step_end_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(t_dev);
    MI_CHECK_LAUNCH();
    return MI_OK;
}

extern "C" int mi_add_noise(const float* L0, const float* x0, const int* Z, const float* z_l, const float* z_x,
                            const float* z_a, int B, int N, int A, const float* coef, const int* t_dev, int t_host,
                            float* l_t, float* x_t, float* a_t, float* tar_x, mi_stream_t stream) {
    if (N <= 0 || B <= 0) return MI_OK;
    MI_CHECK_ARG(L0 && x0 && Z && z_l && z_x && z_a && l_t && x_t && a_t && tar_x && A > 0 && coef, "null pointer");
    long long na = (long long)N * A;
    long long n = na > 9LL * B ? na : 9LL * B;
    add_noise_kernel<<<mi_div_up(n, 256), 256, 0, (cudaStream_t)stream>>>(L0, x0, Z, z_l, z_x, z_a, B, N, A, coef, t_dev,
                                                                          t_host, l_t, x_t, a_t, tar_x);
    MI_CHECK_LAUNCH();
    return MI_OK;
}

extern "C" int mi_rl_loss(const float* pred_l, const float* pred_x, const float* pred_a, const float* tgt_l,
                          const float* tgt_x, const float* tgt_a, const float* prior_l, const float* prior_x,
                          const float* prior_a, const int* node_off, int B, int A, float cost_l, float cost_x,
                          float cost_a, const float* w_loss, const float* w_kl, float scale, float* loss, float* kl,
                          float* d_l, float* d_x, float* d_a, float* stats, mi_stream_t stream) {
    if (B <= 0) return MI_OK;
    MI_CHECK_ARG(pred_l && pred_x && pred_a && node_off && A > 0, "null pointer");
    MI_CHECK_ARG((!prior_l) == (!prior_x) && (!prior_l) == (!prior_a), "prior predictions must be all set or all NULL");
    MI_CHECK_ARG((!tgt_l) == (!tgt_x) && (!tgt_l) == (!tgt_a), "targets must be all set or all NULL");
    rl_loss_kernel<<<B, 128, 0, (cudaStream_t)stream>>>(pred_l, pred_x, pred_a, tgt_l, tgt_x, tgt_a, prior_l, prior_x, prior_a,
                                                        node_off, A, cost_l, cost_x, cost_a, w_loss, w_kl, scale, loss, kl,
                                                        d_l, d_x, d_a, stats);
    MI_CHECK_LAUNCH();
    return MI_OK;
}

extern "C" int mi_adam_step(float* p, float* g, float* m, float* v, long long n, double lr, double b1, double b2,
                            double eps, int step, float grad_scale, int zero_grad, mi_stream_t stream) {
    if (n <= 0) return MI_OK;
    MI_CHECK_ARG(p && g && m && v && step >= 1, "bad arguments");
    MI_CHECK_ARG(mi_host_aligned16(p) && mi_host_aligned16(g) && mi_host_aligned16(m) && mi_host_aligned16(v), "buffers must be 16-byte aligned");
    double bc1 = 1.0 - pow(b1, (double)step);
    double bc2 = 1.0 - pow(b2, (double)step);
    float step_size = (float)(lr / bc1);
    float bc2_sqrt = (float)sqrt(bc2);
    adam_kernel<<<mi_div_up(mi_div_up(n, 4), 256), 256, 0, (cudaStream_t)stream>>>(
        p, g, m, v, n, step_size, (float)(1.0 - b1), (float)b2, (float)(1.0 - b2), (float)eps, bc2_sqrt, grad_scale, zero_grad);
    MI_CHECK_LAUNCH();
    return MI_OK;
}

static int philox_launch(bool normal, float* out, long long n, unsigned long long seed, unsigned long long offset,
                         unsigned long long* offset_dev, int advance, mi_stream_t stream) {
    if (n <= 0) return MI_OK;
    MI_CHECK_ARG(out != nullptr, "null pointer");
    long long q = (n + 3) / 4;
    cudaStream_t s = (cudaStream_t)stream;
    if (normal) philox_fill_kernel<true><<<mi_div_up(q, 256), 256, 0, s>>>(out, n, seed, offset, offset_dev);
    else philox_fill_kernel<false><<<mi_div_up(q, 256), 256, 0, s>>>(out, n, seed, offset, offset_dev);
    MI_CHECK_LAUNCH();
    if (offset_dev && advance) {
        advance_offset_kernel<<<1, 1, 0, s>>>(offset_dev, (unsigned long long)q);
        MI_CHECK_LAUNCH();
    }
    return MI_OK;
}
extern "C" int mi_philox_normal(float* out, long long n, unsigned long long seed, unsigned long long offset,
                                unsigned long long* offset_dev, int advance, mi_stream_t stream) {
    return philox_launch(true, out, n, seed, offset, offset_dev, advance, stream);
}
extern "C" int mi_philox_uniform(float* out, long long n, unsigned long long seed, unsigned long long offset,
                                 unsigned long long* offset_dev, int advance, mi_stream_t stream) {
    return philox_launch(false, out, n, seed, offset, offset_dev, advance, stream);
}
// ------------------------------------------------------------------------------------ transposes for the weight gradients
// dW = dY^T X runs on the tensor cores as the forward kernel with both operands transposed to K(=rows)-contiguous form:
//   XT[c][m] = X[m][c]  (fp32, for the A role)           + col_amax[c] = max_m |X[m][c]|
//   XT_hi / XT_lo[c][m] = merged-format fp16 split of s_c X[m][c], s_c = 2^(14 - exponent(col_amax[c]))  (the W role)
// 32 x 32 tiles through shared memory; block (32, 8); pad columns m in [M, ldt) are left untouched (callers zero them once).
__global__ void transpose_amax_kernel(const float* __restrict__ X, int ldx, int M, int C, float* __restrict__ XT, int ldt,
                                      float* __restrict__ col_amax) {
    __shared__ float tile[32][33];
    const int m0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    float mx = 0.f;
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
        const int m = m0 + threadIdx.y + j, c = c0 + threadIdx.x;
        const float v = (m < M && c < C) ? __ldg(X + (long long)m * ldx + c) : 0.f;
        tile[threadIdx.y + j][threadIdx.x] = v;
        mx = fmaxf(mx, fabsf(v));
    }
    if (col_amax) {
        __shared__ float red[8][32];
        red[threadIdx.y][threadIdx.x] = mx;
        __syncthreads();
        if (threadIdx.y == 0) {
#pragma unroll
            for (int j = 1; j < 8; ++j) mx = fmaxf(mx, red[j][threadIdx.x]);
            if (c0 + threadIdx.x < C) atomicMax(reinterpret_cast<unsigned*>(col_amax + c0 + threadIdx.x), __float_as_uint(mx));
        }
    } else {
        __syncthreads();
    }
    if (XT) {
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
            const int c = c0 + threadIdx.y + j, m = m0 + threadIdx.x;
            if (c < C && m < M) XT[(long long)c * ldt + m] = tile[threadIdx.x][threadIdx.y + j];
        }
    }
}
__global__ void transpose_split_kernel(const float* __restrict__ X, int ldx, int M, int C, const float* __restrict__ col_amax,
                                       __half* __restrict__ hi, __half* __restrict__ lo, int ldt,
                                       float* __restrict__ inv_scale) {
    __shared__ float tile[32][33];
    const int m0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
        const int m = m0 + threadIdx.y + j, c = c0 + threadIdx.x;
        tile[threadIdx.y + j][threadIdx.x] = (m < M && c < C) ? __ldg(X + (long long)m * ldx + c) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
        const int c = c0 + threadIdx.y + j, m = m0 + threadIdx.x;
        if (c >= C) continue;
        const float amax = __ldg(col_amax + c);
        int ex = (int)((__float_as_uint(amax) >> 23) & 0xff) - 127;
        if (amax == 0.f || ex > 100) ex = 14;
        ex = max(ex, -100);
        if (m < M) {
            const float x = tile[threadIdx.x][threadIdx.y + j] * __uint_as_float((uint32_t)(127 + 14 - ex) << 23);
            const __half h = __float2half_rn(x);
            hi[(long long)c * ldt + m] = h;
            lo[(long long)c * ldt + m] = __float2half_rn(x - __half2float(h));
        }
        if (m0 == 0 && threadIdx.x == 0) inv_scale[c] = __uint_as_float((uint32_t)(127 - 14 + ex) << 23);
    }
}



extern "C" int mi_transpose_amax(const float* X, int ldx, int M, int C, float* XT, int ldt, float* col_amax,
                                 mi_stream_t stream) {
    if (M <= 0 || C <= 0) return MI_OK;
    MI_CHECK_ARG(X && (XT || col_amax) && ldx >= C && (!XT || ldt >= M), "bad arguments");
    transpose_amax_kernel<<<dim3(mi_div_up(M, 32), mi_div_up(C, 32)), dim3(32, 8), 0, (cudaStream_t)stream>>>(X, ldx, M, C, XT, ldt, col_amax);
    MI_CHECK_LAUNCH();
    return MI_OK;
}
extern "C" int mi_transpose_split(const float* X, int ldx, int M, int C, const float* col_amax, void* hi, void* lo, int ldt,
                                  float* inv_scale, mi_stream_t stream) {
    if (M <= 0 || C <= 0) return MI_OK;
    MI_CHECK_ARG(X && col_amax && hi && lo && inv_scale && ldx >= C && ldt >= M, "bad arguments");
    transpose_split_kernel<<<dim3(mi_div_up(M, 32), mi_div_up(C, 32)), dim3(32, 8), 0, (cudaStream_t)stream>>>(
        X, ldx, M, C, col_amax, (__half*)hi, (__half*)lo, ldt, inv_scale);
    MI_CHECK_LAUNCH();
    return MI_OK;
}
