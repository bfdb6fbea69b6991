// Shared helpers for libmatinvent_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>

#include "../../include/matinvent_b200.h"

extern "C" void mi_set_error_(const char* fmt, ...);

#define MI_CHECK_ARG(cond, msg)                                              \
    do {                                                                     \
        if (!(cond)) {                                                       \
            mi_set_error_("%s:%d: bad argument: %s", __FILE__, __LINE__, msg); \
            return MI_ERR_ARG;                                               \
        }                                                                    \
    } while (0)

#define MI_CHECK_LAUNCH()                                                         \
    do {                                                                          \
        cudaError_t e__ = cudaGetLastError();                                     \
        if (e__ != cudaSuccess) {                                                 \
            mi_set_error_("%s:%d: CUDA launch failed: %s", __FILE__, __LINE__,    \
                          cudaGetErrorString(e__));                               \
            return MI_ERR_CUDA;                                                   \
        }                                                                         \
    } while (0)

#define MI_CUDA(call)                                                             \
    do {                                                                          \
        cudaError_t e__ = (call);                                                 \
        if (e__ != cudaSuccess) {                                                 \
            mi_set_error_("%s:%d: %s failed: %s", __FILE__, __LINE__, #call,      \
                          cudaGetErrorString(e__));                               \
            return MI_ERR_CUDA;                                                   \
        }                                                                         \
    } while (0)

static inline int mi_div_up(long long a, long long b) { return (int)((a + b - 1) / b); }

__device__ __forceinline__ float mi_silu(float x) { return x / (1.0f + expf(-x)); }
// d/dx [x * sigmoid(x)] = s * (1 + x * (1 - s))
__device__ __forceinline__ float mi_dsilu(float x) {
    float s = 1.0f / (1.0f + expf(-x));
    return s * (1.0f + x * (1.0f - s));
}
// torch.remainder(a, 1.0) for fp32: fmod, then shift negatives up by one (can return exactly 1.0f
// for tiny negative inputs, like the reference: SURVEY.md Appendix C).
__device__ __forceinline__ float mi_mod1(float a) {
    float r = fmodf(a, 1.0f);
    if (r != 0.0f && r < 0.0f) r += 1.0f;
    return r;
}

__device__ __forceinline__ float mi_warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ bool mi_aligned16(const void* p) { return (((uintptr_t)p) & 15u) == 0; }
static inline bool mi_host_aligned16(const void* p) { return (((uintptr_t)p) & 15u) == 0; }
