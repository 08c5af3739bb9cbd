// Fused residual add + LayerNorm of the teacher-forced forward (row a2, reference timing.py:57-58;
// upstream ResidualAttentionBlock: `x = x + attn(ln(x))`): y = x + h, n = LayerNorm(y) * gamma + beta in
// one pass.  torch runs these as an elementwise add (read 2, write 1) and a LayerNorm (read 1, write 1)
// whose fp32 kernel reaches ~35 % of the HBM roofline on (48000, 1024); fused it is read 2 / write 2 at
// streaming speed.  One warp per row, the row lives in registers (kVec float4 per lane), statistics by
// shuffle trees: mean first, then the variance of the centred values (two-pass, fp32).
#include "common.cuh"

namespace wca {

template <int kVec>
__global__ void __launch_bounds__(256) add_layernorm_kernel(const float *__restrict__ x, const float *__restrict__ h,
                                                            const float *__restrict__ gamma, const float *__restrict__ beta,
                                                            float *__restrict__ y, float *__restrict__ n, int64_t rows,
                                                            float eps) {
    constexpr int kWidth = kVec * 128;
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const float4 *xr = reinterpret_cast<const float4 *>(x + row * kWidth);
    const float4 *hr = h ? reinterpret_cast<const float4 *>(h + row * kWidth) : nullptr;
    float4 v[kVec];
#pragma unroll
    for (int k = 0; k < kVec; ++k) v[k] = xr[k * 32 + lane];  // all loads in flight
    if (hr) {
        float4 a[kVec];
#pragma unroll
        for (int k = 0; k < kVec; ++k) a[k] = hr[k * 32 + lane];
#pragma unroll
        for (int k = 0; k < kVec; ++k) {
            v[k].x += a[k].x;
            v[k].y += a[k].y;
            v[k].z += a[k].z;
            v[k].w += a[k].w;
        }
        if (y) {
            float4 *yr = reinterpret_cast<float4 *>(y + row * kWidth);
#pragma unroll
            for (int k = 0; k < kVec; ++k) yr[k * 32 + lane] = v[k];
        }
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < kVec; ++k) s += (v[k].x + v[k].y) + (v[k].z + v[k].w);
    const float mean = warp_sum(s) * (1.f / kWidth);
    float q = 0.f;
#pragma unroll
    for (int k = 0; k < kVec; ++k) {
        const float dx = v[k].x - mean, dy = v[k].y - mean, dz = v[k].z - mean, dw = v[k].w - mean;
        q += (dx * dx + dy * dy) + (dz * dz + dw * dw);
    }
    const float rstd = 1.f / sqrtf(warp_sum(q) * (1.f / kWidth) + eps);
    const float4 *g4 = reinterpret_cast<const float4 *>(gamma);
    const float4 *b4 = reinterpret_cast<const float4 *>(beta);
    float4 *nr = reinterpret_cast<float4 *>(n + row * kWidth);
#pragma unroll
    for (int k = 0; k < kVec; ++k) {
        const float4 g = g4[k * 32 + lane], b = b4[k * 32 + lane];
        float4 o;
        o.x = (v[k].x - mean) * rstd * g.x + b.x;
        o.y = (v[k].y - mean) * rstd * g.y + b.y;
        o.z = (v[k].z - mean) * rstd * g.z + b.z;
        o.w = (v[k].w - mean) * rstd * g.w + b.w;
        nr[k * 32 + lane] = o;
    }
}

template <int kVec>
static int launch_vec(const float *x, const float *h, const float *gamma, const float *beta, float *y, float *n, int64_t rows,
                      float eps, cudaStream_t stream) {
    const int warps = 8;
    const int64_t blocks = (rows + warps - 1) / warps;
    add_layernorm_kernel<kVec><<<(unsigned)blocks, warps * 32, 0, stream>>>(x, h, gamma, beta, y, n, rows, eps);
    WCA_LAUNCH_CHECK("add_layernorm_kernel");
    return WCA_OK;
}

int launch_add_layernorm(const float *x, const float *h, const float *gamma, const float *beta, float *y, float *n,
                         int64_t rows, int width, float eps, cudaStream_t stream) {
    switch (width / 128) {
        case 1: return launch_vec<1>(x, h, gamma, beta, y, n, rows, eps, stream);
        case 2: return launch_vec<2>(x, h, gamma, beta, y, n, rows, eps, stream);
        case 3: return launch_vec<3>(x, h, gamma, beta, y, n, rows, eps, stream);
        case 4: return launch_vec<4>(x, h, gamma, beta, y, n, rows, eps, stream);
        case 6: return launch_vec<6>(x, h, gamma, beta, y, n, rows, eps, stream);
        case 8: return launch_vec<8>(x, h, gamma, beta, y, n, rows, eps, stream);
        case 10: return launch_vec<10>(x, h, gamma, beta, y, n, rows, eps, stream);
        default:
            set_error("wca_add_layernorm: width=%d has no instantiation (128, 256, 384, 512, 768, 1024, 1280)", width);
            return WCA_ERR_UNSUPPORTED;
    }
}

}  // namespace wca
