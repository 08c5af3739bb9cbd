// Fused median filter -> *qk_scale -> softmax along frames (north-star kernel 2).
// Replaces reference timing.py:64-66 and upstream whisper.timing.median_filter (whose
// CUDA branch is a Triton kernel; this is the sm_100a replacement).
//
// One warp per row, the row staged once in shared memory, HBM traffic = one read of the
// used F columns + one write: 8 bytes per map element.
#include "medfilt_softmax.cuh"

namespace wca {

template <int W>
__global__ void __launch_bounds__(256) medfilt_softmax_rows_kernel(const float *__restrict__ in, int64_t n_rows,
                                                                   int64_t ld_in, int F, int width, float qk_scale,
                                                                   float *out, int64_t per_warp_floats) {
    extern __shared__ float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int warps = blockDim.x >> 5;
    float *padded = smem + (int64_t)warp * per_warp_floats;
    float *filtered = padded + F + 2 * (width / 2);
    for (int64_t row = (int64_t)blockIdx.x * warps + warp; row < n_rows; row += (int64_t)gridDim.x * warps)
        filter_softmax_row<W>(in + row * ld_in, out + row * (int64_t)F, F, width, qk_scale, padded, filtered, lane);
}

// In-place variant over a batch: utterance u owns rows_per_frame_block = n_heads*T_u rows
// of F_u floats starting at ws + ws_off.
template <int W>
__global__ void __launch_bounds__(256) medfilt_softmax_batched_kernel(float *ws, const wca_utt_t *__restrict__ utts,
                                                                      int n_heads, int width, float qk_scale,
                                                                      int64_t per_warp_floats) {
    extern __shared__ float smem[];
    const wca_utt_t u = utts[blockIdx.y];
    const int F = u.n_frames;
    const int64_t n_rows = (int64_t)n_heads * u.n_tokens;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int warps = blockDim.x >> 5;
    float *padded = smem + (int64_t)warp * per_warp_floats;
    float *filtered = padded + F + 2 * (width / 2);
    float *base = ws + u.ws_off;
    for (int64_t row = (int64_t)blockIdx.x * warps + warp; row < n_rows; row += (int64_t)gridDim.x * warps)
        filter_softmax_row<W>(base + row * F, base + row * F, F, width, qk_scale, padded, filtered, lane);
}

static int pick_warps(int F, int width, int64_t *per_warp_floats, size_t *smem_bytes) {
    *per_warp_floats = row_smem_floats(F, width);
    int warps = 8;
    while (warps > 1 && (size_t)warps * (*per_warp_floats) * sizeof(float) > 200u * 1024u) warps >>= 1;
    *smem_bytes = (size_t)warps * (*per_warp_floats) * sizeof(float);
    return warps;
}

template <typename K>
static int raise_smem(K kernel, size_t bytes) {
    if (bytes > 48u * 1024u)
        WCA_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return WCA_OK;
}

int launch_medfilt_softmax_rows(const float *d_in, int64_t n_rows, int64_t ld_in, int F, int width, float qk_scale,
                                float *d_out, int sm_count, cudaStream_t stream) {
    int64_t pwf;
    size_t smem;
    const int warps = pick_warps(F, width, &pwf, &smem);
    if (smem > 227u * 1024u) {
        set_error("wca_medfilt_softmax: n_frames=%d does not fit in shared memory", F);
        return WCA_ERR_UNSUPPORTED;
    }
    const int64_t want = (n_rows + warps - 1) / warps;
    const int64_t cap = (int64_t)sm_count * 16;
    const dim3 grid((unsigned)(want < cap ? want : cap)), block(32 * warps);
#define WCA_GO(W)                                                                                              \
    do {                                                                                                       \
        int rc = raise_smem(medfilt_softmax_rows_kernel<W>, smem);                                             \
        if (rc) return rc;                                                                                     \
        medfilt_softmax_rows_kernel<W><<<grid, block, smem, stream>>>(d_in, n_rows, ld_in, F, width, qk_scale, \
                                                                      d_out, pwf);                             \
    } while (0)
    switch (width) {
        case 1: WCA_GO(1); break;
        case 3: WCA_GO(3); break;
        case 5: WCA_GO(5); break;
        case 7: WCA_GO(7); break;
        default: WCA_GO(0); break;
    }
#undef WCA_GO
    WCA_LAUNCH_CHECK("medfilt_softmax_rows_kernel");
    return WCA_OK;
}

int launch_medfilt_softmax_batched(float *d_ws, const wca_utt_t *d_utts, int n_utts, int n_heads, int max_tokens,
                                   int max_frames, int width, float qk_scale, int sm_count, cudaStream_t stream) {
    int64_t pwf;
    size_t smem;
    const int warps = pick_warps(max_frames, width, &pwf, &smem);
    if (smem > 227u * 1024u) {
        set_error("medfilt_softmax: max_frames=%d does not fit in shared memory", max_frames);
        return WCA_ERR_UNSUPPORTED;
    }
    const int64_t rows = (int64_t)n_heads * max_tokens;
    int64_t gx = (rows + warps - 1) / warps;
    const int64_t cap = ((int64_t)sm_count * 16 + n_utts - 1) / n_utts;
    if (gx > cap) gx = cap;
    if (gx < 1) gx = 1;
    const dim3 grid((unsigned)gx, (unsigned)n_utts), block(32 * warps);
#define WCA_GO(W)                                                                                            \
    do {                                                                                                     \
        int rc = raise_smem(medfilt_softmax_batched_kernel<W>, smem);                                        \
        if (rc) return rc;                                                                                   \
        medfilt_softmax_batched_kernel<W><<<grid, block, smem, stream>>>(d_ws, d_utts, n_heads, width,       \
                                                                         qk_scale, pwf);                     \
    } while (0)
    switch (width) {
        case 1: WCA_GO(1); break;
        case 3: WCA_GO(3); break;
        case 5: WCA_GO(5); break;
        case 7: WCA_GO(7); break;
        default: WCA_GO(0); break;
    }
#undef WCA_GO
    WCA_LAUNCH_CHECK("medfilt_softmax_batched_kernel");
    return WCA_OK;
}

}  // namespace wca
