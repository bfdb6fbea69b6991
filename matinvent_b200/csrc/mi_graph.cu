// Periodic neighbour list (edge_style='knn') and the device-resident replay buffer primitives.
#include "mi_common.cuh"

namespace {

constexpr int RG_WARPS = 8;

__device__ __forceinline__ bool cell_earlier(int c) {
    // c = (c0+1)*9 + (c1+1)*3 + (c2+1); lexicographically negative offset (cspnet.py:182-191)
    int c0 = c / 9 - 1, c1 = (c / 3) % 3 - 1, c2 = c % 3 - 1;
    return (c0 < 0) || (c0 == 0 && c1 < 0) || (c0 == 0 && c1 == 0 && c2 < 0);
}

// One CTA per crystal.  Dynamic smem: pos[max_n*3] | offs[27*3] | dist[RG_WARPS][27*max_n] |
// kept[max_n][words] (words = ceil(27*max_n/32)) | thr[RG_WARPS]
__global__ void __launch_bounds__(RG_WARPS * 32) radius_graph_kernel(
    const float* __restrict__ x, const float* __restrict__ L, const int* __restrict__ node_off, int max_n,
    int K, int cap, int* __restrict__ edge_dst, float* __restrict__ cell_off, int* __restrict__ deg,
    int* __restrict__ overflow) {
    extern __shared__ float smem[];
    const int b = blockIdx.x;
    const int n0 = node_off[b], n = node_off[b + 1] - n0;
    const int nq = 27 * n;
    const int words = (27 * max_n + 31) / 32;
    float* pos = smem;
    float* offs = pos + 3 * max_n;
    float* dist = offs + 81;
    unsigned* kept = reinterpret_cast<unsigned*>(dist + RG_WARPS * 27 * max_n);
    float* thr = reinterpret_cast<float*>(kept + max_n * words);
    __shared__ float radius2;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* l = L + 9 * b;

    // cartesian positions: einsum('bi,bij->bj') (cspnet.py:245-246)
    for (int t = tid; t < 3 * n; t += blockDim.x) {
        int a = t / 3, k = t % 3;
        const float* xa = x + 3 * (n0 + a);
        pos[t] = xa[0] * l[k] + xa[1] * l[3 + k] + xa[2] * l[6 + k];
    }
    // image offsets: cell^T @ unit_cell (utils.py:427-431)
    for (int t = tid; t < 81; t += blockDim.x) {
        int c = t / 3, k = t % 3;
        float c0 = (float)(c / 9 - 1), c1 = (float)((c / 3) % 3 - 1), c2 = (float)(c % 3 - 1);
        offs[t] = l[k] * c0 + l[3 + k] * c1 + l[6 + k] * c2;
    }
    if (tid == 0) {
        // radius = min inter-plane distance + 0.01 (utils.py:399-410, 463); `radius` argument ignored
        float a0[3] = {l[0], l[1], l[2]}, a1[3] = {l[3], l[4], l[5]}, a2[3] = {l[6], l[7], l[8]};
        float c23[3] = {a1[1] * a2[2] - a1[2] * a2[1], a1[2] * a2[0] - a1[0] * a2[2], a1[0] * a2[1] - a1[1] * a2[0]};
        float c31[3] = {a2[1] * a0[2] - a2[2] * a0[1], a2[2] * a0[0] - a2[0] * a0[2], a2[0] * a0[1] - a2[1] * a0[0]};
        float c12[3] = {a0[1] * a1[2] - a0[2] * a1[1], a0[2] * a1[0] - a0[0] * a1[2], a0[0] * a1[1] - a0[1] * a1[0]};
        float vol = a0[0] * c23[0] + a0[1] * c23[1] + a0[2] * c23[2];
        auto inv_norm = [&](const float* c) {
            float u = c[0] / vol, v = c[1] / vol, w = c[2] / vol;
            return 1.0f / sqrtf(u * u + v * v + w * w);
        };
        float r = fminf(fminf(inv_norm(c23), inv_norm(c31)), inv_norm(c12)) + 0.01f;
        radius2 = r * r;
    }
    __syncthreads();

    // ---- phase 1: per centre atom, candidate distances and the neighbour cap (utils.py:517-601)
    float* dw = dist + warp * 27 * max_n;
    for (int i = warp; i < n; i += RG_WARPS) {
        int count = 0;
        for (int q0 = 0; q0 < nq; q0 += 32) {
            int q = q0 + lane;
            float d2 = INFINITY;
            if (q < nq) {
                int j = q / 27, c = q - j * 27;
                float dx = pos[3 * i] - (pos[3 * j] + offs[3 * c]);
                float dy = pos[3 * i + 1] - (pos[3 * j + 1] + offs[3 * c + 1]);
                float dz = pos[3 * i + 2] - (pos[3 * j + 2] + offs[3 * c + 2]);
                float v = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
                if (v <= radius2 && v > 0.0001f) d2 = v;
                dw[q] = d2;
            }
            count += __popc(__ballot_sync(0xffffffffu, d2 != INFINITY));
        }
        __syncwarp();
        float th = INFINITY;
        if (count > K) {
            // element of rank K (0-based) in the ascending order = the (K+1)-th nearest
            if (lane == 0) thr[warp] = INFINITY;
            __syncwarp();
            for (int q = lane; q < nq; q += 32) {
                float d = dw[q];
                if (d == INFINITY) continue;
                int rank = 0;
                for (int p = 0; p < nq; ++p) {
                    float e = dw[p];
                    rank += (e < d) || (e == d && p < q);
                }
                if (rank == K) thr[warp] = d + 0.01f;
            }
            __syncwarp();
            th = thr[warp];
        }
        for (int q0 = 0; q0 < nq; q0 += 32) {
            int q = q0 + lane;
            bool k = (q < nq) && (dw[q] != INFINITY) && (dw[q] < th);
            unsigned m = __ballot_sync(0xffffffffu, k);
            if (lane == 0) kept[i * words + (q0 >> 5)] = m;
        }
        __syncwarp();
    }
    __syncthreads();

    // ---- phase 2: symmetric edge list (cspnet.py:159-257), emitted per SOURCE node in a fixed order
    auto is_kept = [&](int i, int q) { return (kept[i * words + (q >> 5)] >> (q & 31)) & 1u; };
    for (int s = warp; s < n; s += RG_WARPS) {
        int total = 0;
        const long long base = (long long)(n0 + s) * cap;
        // (I) centre s, neighbour j < s (or j == s in an "earlier" cell): edge s -> j, offset +c
        // (II) centre i > s (or i == s, earlier cell) holding neighbour s: edge s -> i, offset -c
        const int n1 = (s + 1) * 27;           // q = j*27 + c, j <= s
        const int n2 = (n - s) * 27;           // r = (i - s)*27 + c, i >= s
        for (int e0 = 0; e0 < n1 + n2; e0 += 32) {
            int e = e0 + lane;
            bool ok = false;
            int dst = 0, c = 0;
            float sign = 1.f;
            if (e < n1) {
                int j = e / 27;
                c = e - j * 27;
                ok = is_kept(s, e) && (j < s || cell_earlier(c));
                dst = j;
            } else if (e < n1 + n2) {
                int r = e - n1;
                int i = s + r / 27;
                c = r % 27;
                ok = is_kept(i, s * 27 + c) && (i > s || cell_earlier(c));
                dst = i;
                sign = -1.f;
            }
            unsigned m = __ballot_sync(0xffffffffu, ok);
            if (ok) {
                int p = total + __popc(m & ((1u << lane) - 1u));
                if (p < cap) {
                    edge_dst[base + p] = n0 + dst;
                    float* co = cell_off + 3 * (base + p);
                    co[0] = sign * (float)(c / 9 - 1);
                    co[1] = sign * (float)((c / 3) % 3 - 1);
                    co[2] = sign * (float)(c % 3 - 1);
                } else {
                    atomicExch(overflow, 1);
                }
            }
            total += __popc(m);
        }
        if (lane == 0) deg[n0 + s] = min(total, cap);
    }
}

// exclusive scan of deg[N] -> ptr[N+1], single CTA of 1024 threads
__global__ void __launch_bounds__(1024) scan_kernel(const int* __restrict__ deg, int N, int* __restrict__ ptr) {
    __shared__ int warp_sums[32];
    __shared__ int carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < N; base += 1024) {
        int i = base + tid;
        int v = (i < N) ? deg[i] : 0;
        int inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) warp_sums[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            int w = warp_sums[lane];
            int winc = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int t = __shfl_up_sync(0xffffffffu, winc, o);
                if (lane >= o) winc += t;
            }
            warp_sums[lane] = winc - w;   // exclusive
        }
        __syncthreads();
        int excl = carry + warp_sums[warp] + inc - v;
        if (i < N) ptr[i] = excl;
        __syncthreads();
        if (tid == 1023) carry = excl + v;
        __syncthreads();
    }
    if (tid == 0) ptr[N] = carry;
}

__global__ void compact_edges_kernel(const int* __restrict__ deg, int N, int cap, const int* __restrict__ dst_pad,
                                     const float* __restrict__ cell_pad, const int* __restrict__ node_graph,
                                     const int* __restrict__ seg_ptr, int* __restrict__ edge_src,
                                     int* __restrict__ edge_dst, int* __restrict__ edge_graph,
                                     float* __restrict__ cell_off, int E_cap) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)N * cap) return;
    int node = (int)(t / cap), k = (int)(t - (long long)node * cap);
    if (k >= deg[node]) return;
    int e = seg_ptr[node] + k;
    if (e >= E_cap) return;
    edge_src[e] = node;
    edge_dst[e] = dst_pad[t];
    edge_graph[e] = node_graph[node];
    cell_off[3 * (long long)e] = cell_pad[3 * t];
    cell_off[3 * (long long)e + 1] = cell_pad[3 * t + 1];
    cell_off[3 * (long long)e + 2] = cell_pad[3 * t + 2];
}

__global__ void dst_hist_kernel(const int* __restrict__ seg_ptr, const int* __restrict__ edge_dst, int N,
                                int* __restrict__ cnt) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= seg_ptr[N]) return;
    atomicAdd(cnt + edge_dst[e], 1);
}
// deterministic fill: one thread per destination node scans nothing — instead each edge finds its slot by
// counting earlier edges with the same destination inside the same crystal's contiguous edge range.
__global__ void dst_fill_kernel(const int* __restrict__ seg_ptr, const int* __restrict__ edge_dst, int N,
                                const int* __restrict__ dst_ptr, int* __restrict__ cursor,
                                int* __restrict__ dst_perm) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= seg_ptr[N]) return;
    int d = edge_dst[e];
    int slot = atomicAdd(cursor + d, 1);
    dst_perm[dst_ptr[d] + slot] = e;
}
// restore a deterministic order inside every destination segment (ascending edge id)
__global__ void dst_sort_kernel(const int* __restrict__ dst_ptr, int N, int* __restrict__ dst_perm) {
    int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= N) return;
    int b = dst_ptr[d], e = dst_ptr[d + 1];
    for (int i = b + 1; i < e; ++i) {
        int v = dst_perm[i], j = i - 1;
        while (j >= b && dst_perm[j] > v) { dst_perm[j + 1] = dst_perm[j]; --j; }
        dst_perm[j + 1] = v;
    }
}

// ------------------------------------------------------------------------------------ replay buffer
__device__ __forceinline__ unsigned long long mix64(unsigned long long z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
// one warp per crystal: histogram over Z, gcd of the counts, hash of (Z, count/gcd) pairs
__global__ void composition_key_kernel(const int* __restrict__ Z, const int* __restrict__ node_off, int B,
                                       unsigned long long* __restrict__ keys) {
    __shared__ int hist[8][128];
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int b = blockIdx.x * 8 + warp;
    if (b >= B) return;
    int* h = hist[warp];
    for (int k = lane; k < 128; k += 32) h[k] = 0;
    __syncwarp();
    int n0 = node_off[b], n1 = node_off[b + 1];
    for (int i = n0 + lane; i < n1; i += 32) {
        int z = Z[i];
        if (z >= 0 && z < 128) atomicAdd(h + z, 1);
    }
    __syncwarp();
    if (lane == 0) {
        int g = 0;
        for (int k = 0; k < 128; ++k) {
            int a = h[k], c = g;
            while (a) { int t = c % a; c = a; a = t; }
            g = c;
        }
        if (g == 0) g = 1;
        unsigned long long key = 0x243F6A8885A308D3ull;
        for (int k = 0; k < 128; ++k)
            if (h[k]) key = mix64(key ^ mix64(((unsigned long long)k << 32) | (unsigned)(h[k] / g)));
        keys[b] = key;
    }
}

// single CTA (1024 threads): stable sort by reward desc, first-occurrence dedupe by key, head, cutoff
__global__ void __launch_bounds__(1024) replay_select_kernel(const unsigned long long* __restrict__ keys,
                                                             const double* __restrict__ rewards, int n, int npow2,
                                                             int buffer_size, double cutoff, int* __restrict__ out_idx,
                                                             int* __restrict__ out_count) {
    extern __shared__ __align__(16) unsigned char sm_raw[];
    // rewards are compared in float64 like the reference's pandas column (near-equal rewards and the strict cutoff test
    // must not depend on an fp32 rounding)
    double* r = reinterpret_cast<double*>(sm_raw);          // [npow2]
    int* id = reinterpret_cast<int*>(r + npow2);            // [npow2]
    const int tid = threadIdx.x;
    for (int i = tid; i < npow2; i += 1024) {
        r[i] = (i < n) ? rewards[i] : -(double)INFINITY;
        id[i] = (i < n) ? i : 0x7fffffff;
    }
    __syncthreads();
    // "a before b" : reward desc, NaN last, original index asc (stable)
    auto before = [&](double ra, int ia, double rb, int ib) {
        bool pa = ia == 0x7fffffff, pb = ib == 0x7fffffff;   // padding always last
        if (pa != pb) return pb;
        bool na = ra != ra, nb = rb != rb;
        if (na != nb) return nb;
        if (!na && ra != rb) return ra > rb;
        return ia < ib;
    };
    for (int k = 2; k <= npow2; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < npow2; i += 1024) {
                int p = i ^ j;
                if (p > i) {
                    bool up = (i & k) == 0;
                    bool sw = up ? before(r[p], id[p], r[i], id[i]) : before(r[i], id[i], r[p], id[p]);
                    if (sw) {
                        double tr = r[i]; r[i] = r[p]; r[p] = tr;
                        int ti = id[i]; id[i] = id[p]; id[p] = ti;
                    }
                }
            }
            __syncthreads();
        }
    // dedupe: position p survives if no earlier position holds the same key (drop_duplicates keep='first')
    // reuse r[] as the keep flag afterwards, so read rewards back from global.
    for (int p = tid; p < npow2; p += 1024) {
        double keep = 0.0;
        if (p < n) {
            unsigned long long kp = keys[id[p]];
            keep = 1.0;
            for (int q = 0; q < p; ++q)
                if (keys[id[q]] == kp) { keep = 0.0; break; }
        }
        r[p] = keep;   // each thread only touches its own slots of r[] here
    }
    __syncthreads();
    if (tid == 0) {
        // head(buffer_size) of the survivors, then reward > cutoff (replay_buffer.py:60-71); n is small
        int rank = 0, m = 0;
        for (int p = 0; p < n; ++p) {
            if (r[p] == 0.0) continue;
            if (rank < buffer_size && rewards[id[p]] > cutoff) out_idx[m++] = id[p];
            ++rank;
        }
        *out_count = m;
    }
}

}  // namespace

extern "C" int mi_radius_graph_pbc(const float* x, const float* L, const int* node_off, int B, int N, int max_n,
                                   int max_neighbors, int cap, int* edge_dst, float* cell_off, int* deg,
                                   int* overflow, mi_stream_t stream) {
    if (B <= 0) return MI_OK;
    MI_CHECK_ARG(x && L && node_off && edge_dst && cell_off && deg && overflow, "null pointer");
    MI_CHECK_ARG(max_n >= 1 && max_n <= 128 && cap >= 1 && max_neighbors >= 1, "bad sizes");
    int words = (27 * max_n + 31) / 32;
    size_t smem = sizeof(float) * (3 * max_n + 81 + RG_WARPS * 27 * max_n + RG_WARPS) + sizeof(unsigned) * max_n * words;
    if (smem > 48 * 1024) {
        MI_CHECK_ARG(smem <= 227 * 1024, "crystal too large for the neighbour-list kernel");
        MI_CUDA(cudaFuncSetAttribute(radius_graph_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    cudaStream_t s = (cudaStream_t)stream;
    MI_CUDA(cudaMemsetAsync(overflow, 0, sizeof(int), s));
    radius_graph_kernel<<<B, RG_WARPS * 32, smem, s>>>(x, L, node_off, max_n, max_neighbors, cap, edge_dst, cell_off, deg, overflow);
    MI_CHECK_LAUNCH();
    return MI_OK;
}

extern "C" int mi_compact_edges(const int* deg, int N, int cap, const int* edge_dst_pad, const float* cell_pad,
                                const int* node_graph, int* seg_ptr, int* edge_src, int* edge_dst, int* edge_graph,
                                float* cell_off, int E_cap, mi_stream_t stream) {
    if (N <= 0) return MI_OK;
    MI_CHECK_ARG(deg && edge_dst_pad && cell_pad && node_graph && seg_ptr && edge_src && edge_dst && edge_graph && cell_off, "null pointer");
    cudaStream_t s = (cudaStream_t)stream;
    scan_kernel<<<1, 1024, 0, s>>>(deg, N, seg_ptr);
    MI_CHECK_LAUNCH();
    compact_edges_kernel<<<mi_div_up((long long)N * cap, 256), 256, 0, s>>>(deg, N, cap, edge_dst_pad, cell_pad, node_graph, seg_ptr,
                                                                            edge_src, edge_dst, edge_graph, cell_off, E_cap);
    MI_CHECK_LAUNCH();
    return MI_OK;
}

extern "C" int mi_build_dst_csr(const int* seg_ptr, const int* edge_dst, int N, int E_cap, int* dst_ptr, int* dst_perm,
                                int* work, mi_stream_t stream) {
    if (N <= 0 || E_cap <= 0) return MI_OK;
    MI_CHECK_ARG(seg_ptr && edge_dst && dst_ptr && dst_perm && work, "null pointer");
    cudaStream_t s = (cudaStream_t)stream;
    MI_CUDA(cudaMemsetAsync(work, 0, sizeof(int) * (size_t)(N + 1), s));
    dst_hist_kernel<<<mi_div_up(E_cap, 256), 256, 0, s>>>(seg_ptr, edge_dst, N, work);
    MI_CHECK_LAUNCH();
    scan_kernel<<<1, 1024, 0, s>>>(work, N, dst_ptr);
    MI_CHECK_LAUNCH();
    MI_CUDA(cudaMemsetAsync(work, 0, sizeof(int) * (size_t)(N + 1), s));
    dst_fill_kernel<<<mi_div_up(E_cap, 256), 256, 0, s>>>(seg_ptr, edge_dst, N, dst_ptr, work, dst_perm);
    MI_CHECK_LAUNCH();
    dst_sort_kernel<<<mi_div_up(N, 128), 128, 0, s>>>(dst_ptr, N, dst_perm);
    MI_CHECK_LAUNCH();
    return MI_OK;
}

extern "C" int mi_replay_select(const unsigned long long* keys, const double* rewards, int n, int buffer_size,
                                double cutoff, int* out_idx, int* out_count, mi_stream_t stream) {
    MI_CHECK_ARG(out_count != nullptr, "null pointer");
    cudaStream_t s = (cudaStream_t)stream;
    if (n <= 0) {
        MI_CUDA(cudaMemsetAsync(out_count, 0, sizeof(int), s));
        return MI_OK;
    }
    MI_CHECK_ARG(keys && rewards && out_idx && n <= 16384, "null pointer or n > 16384");
    int npow2 = 2;
    while (npow2 < n) npow2 <<= 1;
    size_t smem = (size_t)npow2 * 12;
    if (smem > 48 * 1024)
        MI_CUDA(cudaFuncSetAttribute(replay_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    replay_select_kernel<<<1, 1024, smem, s>>>(keys, rewards, n, npow2, buffer_size, cutoff, out_idx, out_count);
    MI_CHECK_LAUNCH();
    return MI_OK;
}

extern "C" int mi_composition_key(const int* Z, const int* node_off, int B, unsigned long long* keys,
                                  mi_stream_t stream) {
    if (B <= 0) return MI_OK;
    MI_CHECK_ARG(Z && node_off && keys, "null pointer");
    composition_key_kernel<<<mi_div_up(B, 8), 256, 0, (cudaStream_t)stream>>>(Z, node_off, B, keys);
    MI_CHECK_LAUNCH();
    return MI_OK;
}
