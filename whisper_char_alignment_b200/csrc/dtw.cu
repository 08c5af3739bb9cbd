// Batched DTW + on-device backtrace + word-boundary extraction (north-star kernel 4).
// Replaces reference timing.py:102-113: the `.cpu()` round trip, upstream
// whisper.timing.dtw_cpu/backtrace (numba, single core) and the numpy jump extraction.
//
// One CTA per problem, thread r owns text row r+1 of the (N+1) x (M+1) cost table and
// walks it left to right; all rows advance together along anti-diagonals (step s = i + j),
// so each step needs only the neighbour's previous value, exchanged through a
// double-buffered shared-memory line.  The 2-bit trace stays in shared memory (166 KB for
// the largest legal Whisper problem, 445 x 1500), the backtrace runs on the device, and
// only N jump frames / W word times ever leave the SM.
//
// Bit-exactness contract (must match dtw_cpu): fp32 round-to-nearest add, no FMA; the
// diagonal wins only if strictly smaller than both others, then the text step only if
// strictly smaller than both others, otherwise the time step -- so ties and NaN go to
// code 2.  Border rule of backtrace: column 0 -> code 1, row 0 -> code 2.
#include "common.cuh"

namespace wca {

constexpr int kCellsPerWord = 16;  // 2-bit codes

__host__ __device__ inline int trace_words_per_row(int M) {
    int w = (M + kCellsPerWord - 1) / kCellsPerWord;
    return w | 1;  // odd stride: neighbouring rows land in different banks
}

struct DtwLaunch {
    const float *matrix;
    const wca_utt_t *utts;
    int negate;
    int32_t *path_text, *path_time, *path_len, *jump_frames;
    const int32_t *word_bounds;
    double *start_times, *end_times;
    uint32_t *trace_ws;       // global fallback, trace_ws_stride words per problem
    int64_t trace_ws_stride;  // 0 => trace lives in shared memory
};

template <bool kOneWarp>
__device__ __forceinline__ void step_barrier() {
    if constexpr (kOneWarp) __syncwarp();
    else __syncthreads();
}

template <bool kOneWarp>
__global__ void __launch_bounds__(1024) dtw_align_kernel(const DtwLaunch p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const wca_utt_t u = p.utts[blockIdx.x];
    const int N = u.row_end - u.row_begin;
    const int M = u.n_frames;
    const int nthr = blockDim.x;
    const int tid = threadIdx.x;

    float *line = reinterpret_cast<float *>(smem_raw);                 // [2][nthr + 1]
    int32_t *jump_s = reinterpret_cast<int32_t *>(line + 2 * (nthr + 1));  // [nthr]
    uint32_t *trace = p.trace_ws_stride ? p.trace_ws + (int64_t)blockIdx.x * p.trace_ws_stride
                                        : reinterpret_cast<uint32_t *>(jump_s + nthr);
    const int wpr = trace_words_per_row(M);

    if (N <= 0 || M <= 0) {
        if (tid == 0 && p.path_len) p.path_len[blockIdx.x] = 0;
        return;
    }

    // ---- forward sweep ---------------------------------------------------------
    const int i = tid + 1;  // table row
    const bool row_live = i <= N;
    const float *xrow = p.matrix + u.matrix_off + (int64_t)(row_live ? i - 1 : 0) * M;
    for (int k = tid; k < 2 * (nthr + 1); k += nthr) line[k] = INFINITY;  // row 0 border and idle slots
    jump_s[tid] = -1;
    float left = INFINITY;                         // cost[i][j-1], starts at cost[i][0]
    float upleft = (i == 1) ? 0.f : INFINITY;      // cost[i-1][j-1], starts at cost[i-1][0]
    uint32_t tw = 0;
    step_barrier<kOneWarp>();

    auto load_x = [&](int s) -> float {
        const int j = s - i;
        if (!row_live || j < 1 || j > M) return 0.f;
        const float v = xrow[j - 1];
        return p.negate ? -v : v;
    };

    const int s_last = N + M;
    float xc[4], xn[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) xc[q] = load_x(2 + q);
    for (int s0 = 2; s0 <= s_last; s0 += 4) {
#pragma unroll
        for (int q = 0; q < 4; ++q) xn[q] = load_x(s0 + 4 + q);  // in flight while this group runs
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int s = s0 + q;
            if (s <= s_last) {  // uniform across the CTA
                const int j = s - i;
                if (row_live && j >= 1 && j <= M) {
                    const float up = line[((s - 1) & 1) * (nthr + 1) + (i - 1)];  // cost[i-1][j]
                    const float c0 = upleft, c1 = up, c2 = left;
                    float c;
                    uint32_t code;
                    if (c0 < c1 && c0 < c2) {
                        c = c0;
                        code = 0u;
                    } else if (c1 < c0 && c1 < c2) {
                        c = c1;
                        code = 1u;
                    } else {
                        c = c2;
                        code = 2u;
                    }
                    const float cost = __fadd_rn(xc[q], c);
                    line[(s & 1) * (nthr + 1) + i] = cost;
                    upleft = up;
                    left = cost;
                    const int cell = (j - 1) & (kCellsPerWord - 1);
                    tw |= code << (2 * cell);
                    if (cell == kCellsPerWord - 1 || j == M) {
                        trace[(int64_t)(i - 1) * wpr + ((j - 1) >> 4)] = tw;
                        tw = 0;
                    }
                }
                step_barrier<kOneWarp>();
            }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) xc[q] = xn[q];
    }
    __syncthreads();  // trace complete (global fallback: same-CTA visibility is enough)

    // ---- backtrace (sequential by nature) ------------------------------------------
    if (tid == 0) {
        const int cap = N + M;
        int32_t *pt = p.path_text ? p.path_text + u.path_off : nullptr;
        int32_t *pj = p.path_time ? p.path_time + u.path_off : nullptr;
        int bi = N, bj = M, pos = cap;
        while (bi > 0 || bj > 0) {
            --pos;
            if (pt) {
                pt[pos] = bi - 1;
                pj[pos] = bj - 1;
            }
            uint32_t code;
            if (bj == 0) code = 1u;
            else if (bi == 0) code = 2u;
            else code = (trace[(int64_t)(bi - 1) * wpr + ((bj - 1) >> 4)] >> (2 * ((bj - 1) & 15))) & 3u;
            // first path point of a text row: the step out of it changes the row (or ends the walk)
            if ((code != 2u || (bi == 0 && bj == 1)) && bi >= 1) jump_s[bi - 1] = bj - 1;
            if (code == 0u) {
                --bi;
                --bj;
            } else if (code == 1u) {
                --bi;
            } else {
                --bj;
            }
        }
        if (p.path_len) p.path_len[blockIdx.x] = cap - pos;
    }
    __syncthreads();

    // ---- jump frames and word times (timing.py:110-113) ----------------------------
    if (p.jump_frames)
        for (int r = tid; r < N; r += nthr) p.jump_frames[u.jump_off + r] = jump_s[r];
    if (p.word_bounds && p.start_times && p.end_times) {
        const int32_t *wb = p.word_bounds + u.word_off;
        for (int w = tid; w < u.n_words; w += nthr) {
            const int a = wb[w], b = wb[w + 1];
            const double nan = __longlong_as_double(0x7ff8000000000000ll);
            p.start_times[u.word_off + w] = (a >= 0 && a < N) ? (double)jump_s[a] / WCA_TOKENS_PER_SECOND : nan;
            p.end_times[u.word_off + w] = (b >= 0 && b < N) ? (double)jump_s[b] / WCA_TOKENS_PER_SECOND : nan;
        }
    }
}

static size_t dtw_fixed_smem(int nthr) { return (size_t)(2 * (nthr + 1) + nthr) * 4u; }
static size_t dtw_trace_bytes(int max_rows, int max_frames) {
    return (size_t)max_rows * trace_words_per_row(max_frames) * 4u;
}
constexpr size_t kSmemBudget = 227u * 1024u;

int64_t dtw_workspace_bytes(int n_utts, int max_rows, int max_frames) {
    const int nthr = ((max_rows + 31) / 32) * 32;
    if (dtw_fixed_smem(nthr) + dtw_trace_bytes(max_rows, max_frames) <= kSmemBudget) return 0;
    return (int64_t)n_utts * (int64_t)dtw_trace_bytes(max_rows, max_frames);
}

int launch_dtw_align(const float *d_matrix, const wca_utt_t *d_utts, int n_utts, int max_rows, int max_frames,
                     int negate, int32_t *d_path_text, int32_t *d_path_time, int32_t *d_path_len,
                     int32_t *d_jump_frames, const int32_t *d_word_bounds, double *d_start_times,
                     double *d_end_times, void *d_trace_ws, int64_t trace_ws_bytes, cudaStream_t stream) {
    if (max_rows > 1024) {
        set_error("wca_dtw_align: %d text rows exceed the 1024 rows one CTA sweeps (Whisper caps T at 448)", max_rows);
        return WCA_ERR_UNSUPPORTED;
    }
    const int nthr = max_rows <= 0 ? 32 : ((max_rows + 31) / 32) * 32;
    DtwLaunch p;
    p.matrix = d_matrix;
    p.utts = d_utts;
    p.negate = negate;
    p.path_text = d_path_text;
    p.path_time = d_path_time;
    p.path_len = d_path_len;
    p.jump_frames = d_jump_frames;
    p.word_bounds = d_word_bounds;
    p.start_times = d_start_times;
    p.end_times = d_end_times;
    p.trace_ws = nullptr;
    p.trace_ws_stride = 0;
    size_t smem = dtw_fixed_smem(nthr);
    const size_t trace_bytes = dtw_trace_bytes(max_rows > 0 ? max_rows : 1, max_frames > 0 ? max_frames : 1);
    if (smem + trace_bytes <= kSmemBudget) {
        smem += trace_bytes;
    } else {
        const int64_t need = (int64_t)n_utts * (int64_t)trace_bytes;
        if (!d_trace_ws || trace_ws_bytes < need) {
            set_error("wca_dtw_align: trace workspace of %lld bytes required, %lld given", (long long)need,
                      (long long)trace_ws_bytes);
            return WCA_ERR_INVALID;
        }
        p.trace_ws = static_cast<uint32_t *>(d_trace_ws);
        p.trace_ws_stride = (int64_t)(trace_bytes / 4u);
    }
    if (nthr == 32) {
        if (smem > 48u * 1024u)
            WCA_CUDA(cudaFuncSetAttribute(dtw_align_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dtw_align_kernel<true><<<n_utts, nthr, smem, stream>>>(p);
    } else {
        if (smem > 48u * 1024u)
            WCA_CUDA(cudaFuncSetAttribute(dtw_align_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dtw_align_kernel<false><<<n_utts, nthr, smem, stream>>>(p);
    }
    WCA_LAUNCH_CHECK("dtw_align_kernel");
    return WCA_OK;
}

}  // namespace wca
