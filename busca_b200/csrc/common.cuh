// Shared helpers for the sm_100a kernels of libbusca_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <math_constants.h>

#define PATCH_H 384
#define PATCH_W 128
#define PATCH_C 3
#define PATCH_BYTES (PATCH_H * PATCH_W * PATCH_C)
#define EMB_DIM 512

// token-level constants of busca/encodings.py:11 (max_temp_dist, max_distance_dist, max_size_dist)
#define PE_MAX_T 30
#define PE_MAX_XY 105
#define PE_MAX_SIZE 105
#define PE_CH 172          // channels per axis of PositionalEncoding3D(512): ceil(512/6)*2
#define PE_CH_T 168        // 512 - 2*172

static __device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
static __device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// Activation storage types of the ReID path: float (fp32 mode) or __nv_bfloat16 (bf16 mode).
template <typename T> struct ActIO;
template <> struct ActIO<float> {
    static __device__ __forceinline__ float ld(const float *p) { return *p; }
    static __device__ __forceinline__ void st(float *p, float v) { *p = v; }
};
template <> struct ActIO<__nv_bfloat16> {
    static __device__ __forceinline__ float ld(const __nv_bfloat16 *p) { return __bfloat162float(*p); }
    static __device__ __forceinline__ void st(__nv_bfloat16 *p, float v) { *p = __float2bfloat16_rn(v); }
};
