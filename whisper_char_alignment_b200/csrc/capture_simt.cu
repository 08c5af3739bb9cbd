// CUDA-core cross-attention capture: fp32 logits q.k for every (utterance, layer, head),
// trimmed to the utterance's max_frames while being produced (reference timing.py:50-64 +
// upstream MultiHeadAttention.qkv_attention with SDPA disabled).
//
// This is the exact-fp32 cross-check of the tcgen05 kernel (capture_tc.cu) and the
// fallback for head widths it does not cover; selected with WCA_CAPTURE_FORCE_SIMT.
#include "common.cuh"

namespace wca {

constexpr int kTokTile = 32;    // tokens per CTA
constexpr int kFrmTile = 64;    // frames per inner chunk
constexpr int kKStride = kHeadDim + 1;  // +1 float: lane<->frame reads hit distinct banks

__global__ void __launch_bounds__(256) capture_logits_simt_kernel(const __grid_constant__ LayerPtrs ptrs,
                                                                  const wca_utt_t *__restrict__ utts, int n_heads,
                                                                  int64_t ld_q, int64_t ld_k, float scale,
                                                                  float *__restrict__ ws) {
    __shared__ float q_s[kTokTile][kHeadDim];
    __shared__ float k_s[kFrmTile][kKStride];

    const wca_utt_t u = utts[blockIdx.z];
    const int T = u.n_tokens, F = u.n_frames;
    const int t0 = blockIdx.x * kTokTile;
    if (t0 >= T) return;
    const int layer = blockIdx.y / n_heads, head = blockIdx.y % n_heads;
    const float *q = ptrs.q[layer] + u.q_row0 * ld_q + (int64_t)head * kHeadDim;
    const float *k = ptrs.k[layer] + u.k_row0 * ld_k + (int64_t)head * kHeadDim;
    float *out = ws + u.ws_off + (int64_t)blockIdx.y * T * F;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // Q tile, scaled by Dh^-1/4 exactly as the reference scales it before the product.
    for (int i = tid; i < kTokTile * (kHeadDim / 4); i += blockDim.x) {
        const int r = i / (kHeadDim / 4), c4 = i % (kHeadDim / 4);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (t0 + r < T) v = *reinterpret_cast<const float4 *>(q + (int64_t)(t0 + r) * ld_q + 4 * c4);
        q_s[r][4 * c4 + 0] = v.x * scale;
        q_s[r][4 * c4 + 1] = v.y * scale;
        q_s[r][4 * c4 + 2] = v.z * scale;
        q_s[r][4 * c4 + 3] = v.w * scale;
    }

    for (int f0 = 0; f0 < F; f0 += kFrmTile) {
        __syncthreads();  // previous chunk consumed (also orders the Q stores on the first trip)
        for (int i = tid; i < kFrmTile * (kHeadDim / 4); i += blockDim.x) {
            const int r = i / (kHeadDim / 4), c4 = i % (kHeadDim / 4);
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (f0 + r < F) v = *reinterpret_cast<const float4 *>(k + (int64_t)(f0 + r) * ld_k + 4 * c4);
            k_s[r][4 * c4 + 0] = v.x * scale;
            k_s[r][4 * c4 + 1] = v.y * scale;
            k_s[r][4 * c4 + 2] = v.z * scale;
            k_s[r][4 * c4 + 3] = v.w * scale;
        }
        __syncthreads();

        float acc[4][2] = {};
#pragma unroll 16
        for (int c = 0; c < kHeadDim; ++c) {
            const float ka = k_s[lane][c], kb = k_s[lane + 32][c];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const float qv = q_s[4 * warp + r][c];
                acc[r][0] = fmaf(qv, ka, acc[r][0]);
                acc[r][1] = fmaf(qv, kb, acc[r][1]);
            }
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int t = t0 + 4 * warp + r;
            if (t >= T) continue;
            float *row = out + (int64_t)t * F + f0;
            if (f0 + lane < F) row[lane] = acc[r][0];
            if (f0 + lane + 32 < F) row[lane + 32] = acc[r][1];
        }
    }
}

int launch_capture_logits_simt(const float *const *h_q_layers, const float *const *h_k_layers, int n_layers,
                               int n_heads, int64_t ld_q, int64_t ld_k, const wca_utt_t *d_utts, int n_utts,
                               int max_tokens, float *d_ws, cudaStream_t stream) {
    LayerPtrs p;
    for (int l = 0; l < WCA_MAX_LAYERS; ++l) {
        p.q[l] = l < n_layers ? h_q_layers[l] : nullptr;
        p.k[l] = l < n_layers ? h_k_layers[l] : nullptr;
    }
    const float scale = (float)0.35355339059327373;  // 64 ** -0.25, rounded to fp32 like the reference's scalar
    const dim3 grid((max_tokens + kTokTile - 1) / kTokTile, n_layers * n_heads, n_utts), block(256);
    capture_logits_simt_kernel<<<grid, block, 0, stream>>>(p, d_utts, n_heads, ld_q, ld_k, scale, d_ws);
    WCA_LAUNCH_CHECK("capture_logits_simt_kernel");
    return WCA_OK;
}

}  // namespace wca
