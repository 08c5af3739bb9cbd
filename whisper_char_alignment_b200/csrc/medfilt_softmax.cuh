// Row primitive shared by the standalone filter kernel and the CUDA-core capture path:
// one warp turns one row of logits into   softmax( median_w(logits[:F]) * qk_scale ).
// Follows timing.py:64-66 of the reference and whisper.timing.median_filter: reflect
// padding inside the trimmed window, identity filter when F <= w/2.
#pragma once

#include "common.cuh"

namespace wca {

// Shared memory a warp needs for a row of F frames filtered with width w:
// the padded row (F + 2*(w/2)) plus the filtered row (F).
__host__ __device__ inline int64_t row_smem_floats(int F, int width) { return 2ll * F + 2ll * (width / 2); }

template <int W>
__device__ __forceinline__ float median_at(const float *padded, int f, int width) {
    if constexpr (W == 1) {
        return padded[f];
    } else if constexpr (W > 1) {
        float v[W];
#pragma unroll
        for (int i = 0; i < W; ++i) v[i] = padded[f + i];
        return median_regs<W>(v);
    } else {
        return median_window(padded + f, width);
    }
}

// `in` may alias `out` (the row is fully staged in shared memory before any store; the loads therefore use the
// coherent ld.global.cg path, not ld.global.nc).
// padded: F + 2*half floats, filtered: F floats, both private to this warp.
template <int W>
__device__ __forceinline__ void filter_softmax_row(const float *in, float *out, int F, int width, float qk_scale,
                                                   float *padded, float *filtered, int lane) {
    const int half = (F <= width / 2) ? 0 : width / 2;  // identity filter for very short rows
    for (int f = lane; f < F; f += kWarp) padded[half + f] = ld_stream_coherent(in + f);
    __syncwarp();
    // reflect halo (no edge repeat): left i -> x[i+1 .. half], right -> x[F-2 ...]
    for (int i = lane; i < half; i += kWarp) {
        padded[half - 1 - i] = padded[half + 1 + i];
        padded[half + F + i] = padded[half + F - 2 - i];
    }
    __syncwarp();

    float vmax = -INFINITY;
    for (int f = lane; f < F; f += kWarp) {
        float m = (half == 0) ? padded[f] : median_at<W>(padded, f, width);
        m *= qk_scale;
        filtered[f] = m;
        vmax = fmaxf(vmax, m);
    }
    vmax = warp_max(vmax);
    float vsum = 0.f;
    for (int f = lane; f < F; f += kWarp) {
        const float e = expf(filtered[f] - vmax);
        filtered[f] = e;
        vsum += e;
    }
    vsum = warp_sum(vsum);
    for (int f = lane; f < F; f += kWarp) st_stream(out + f, filtered[f] / vsum);
    __syncwarp();
}

}  // namespace wca
