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

// Programmatic dependent launch (PDL): a kernel launched through launch_pdl may become resident while its predecessor in the stream is
// still draining; it runs its private prologue (barrier init, TMEM allocation, tensor-map prefetch), then pdl_wait() blocks until the
// predecessor has COMPLETED and its memory is visible.  Every kernel launched this way calls pdl_wait() before its first access to
// global memory another kernel may write (or still read), and pdl_trigger() at its start so that its own successor may be scheduled
// as SMs free up.  OFF by default (BUSCA_PDL=1 / busca_set_option("pdl", 1) enables it): on a B200 the chain of persistent one-CTA-per-SM
// kernels ran 2 % SLOWER with it (profiles/r02u: 48.0 vs 47.1 ms/frame) - a successor CTA cannot become resident before its
// predecessor's CTA on that SM has exited (shared memory), so only the launch latency is hidden, and griddepcontrol.wait costs more.
static __device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
static __device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
int pdl_enabled();              // api.cu
void pdl_set(int on);
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args &&...args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

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
