// Shared helpers of the sm_100a alignment kernels (host-side error plumbing, warp
// reductions, streaming load/store wrappers).  No torch types anywhere in csrc/.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/wca_b200.h"

namespace wca {

// ---- host-side error plumbing -------------------------------------------------
void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what);
void count_launch();

#define WCA_CHECK_ARG(cond, ...)        \
    do {                                \
        if (!(cond)) {                  \
            ::wca::set_error(__VA_ARGS__); \
            return WCA_ERR_INVALID;     \
        }                               \
    } while (0)

#define WCA_CUDA(call)                                     \
    do {                                                   \
        cudaError_t e__ = (call);                          \
        if (e__ != cudaSuccess) return ::wca::cuda_fail(e__, #call); \
    } while (0)

#define WCA_LAUNCH_CHECK(name)                             \
    do {                                                   \
        cudaError_t e__ = cudaGetLastError();              \
        if (e__ != cudaSuccess) return ::wca::cuda_fail(e__, name); \
        ::wca::count_launch();                             \
    } while (0)

// Per-layer Q / K base pointers, passed to the capture kernels by value.
struct LayerPtrs {
    const float *q[WCA_MAX_LAYERS];
    const float *k[WCA_MAX_LAYERS];
};

constexpr int kHeadDim = 64;   // every published Whisper size has d_head = 64
constexpr int kWarp = 32;

// ---- device helpers -----------------------------------------------------------
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Streaming (read-once / write-once) global accesses: keep them out of L1.
__device__ __forceinline__ float ld_stream(const float *p) {
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
// Same cache policy through the coherent path: for rows a kernel may later overwrite in place (ld.global.nc is only
// defined for memory that stays read-only for the whole kernel).
__device__ __forceinline__ float ld_stream_coherent(const float *p) {
    float v;
    asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float4 ld_stream4(const float4 *p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p));
    return v;
}
__device__ __forceinline__ void st_stream(float *p, float v) {
    asm volatile("st.global.L1::no_allocate.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
__device__ __forceinline__ void st_stream4(float4 *p, float4 v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y),
                 "f"(v.z), "f"(v.w)
                 : "memory");
}

// Index of the reflect-padded sample: position p in [-half, F+half) -> [0, F).
// torch 'reflect' does not repeat the edge sample.  Requires F > half.
__device__ __forceinline__ int reflect_index(int p, int F) {
    if (p < 0) p = -p;
    if (p >= F) p = 2 * (F - 1) - p;
    return p;
}

// Exact median of W values held in registers (compare-exchange network: enough
// bubble passes to settle the upper half, the median is then in place).  min/max
// only, so the result is one of the inputs, bit for bit.
template <int W>
__device__ __forceinline__ float median_regs(float (&v)[W]) {
    constexpr int half = W / 2;
#pragma unroll
    for (int pass = 0; pass <= half; ++pass) {
#pragma unroll
        for (int j = 0; j < W - 1 - pass; ++j) {
            const float lo = fminf(v[j], v[j + 1]);
            const float hi = fmaxf(v[j], v[j + 1]);
            v[j] = lo;
            v[j + 1] = hi;
        }
    }
    return v[half];
}

// Median of an odd-width window read from (shared) memory, any width up to
// WCA_MAX_MEDFILT: rank counting with a stable tie order, so the value returned is
// the same order statistic a stable sort would give.
__device__ __forceinline__ float median_window(const float *win, int width) {
    const int half = width >> 1;
    float result = win[0];
    for (int a = 0; a < width; ++a) {
        const float va = win[a];
        int rank = 0;
        for (int b = 0; b < width; ++b) {
            const float vb = win[b];
            rank += (vb < va) || (vb == va && b < a);
        }
        if (rank == half) result = va;
    }
    return result;
}

}  // namespace wca
