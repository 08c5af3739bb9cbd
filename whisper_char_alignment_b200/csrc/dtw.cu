// Batched DTW + on-device backtrace + word-boundary extraction (north-star kernel 4).
// Replaces reference timing.py:102-113: the `.cpu()` round trip, upstream
// whisper.timing.dtw_cpu/backtrace (numba, single core) and the numpy jump extraction.
//
// One WARP per problem.  Lane l owns a strip of R consecutive text rows (R = ceil(N/32),
// compile-time 1..32) and walks it left to right, one column per step, one step behind lane
// l-1: the anti-diagonal wavefront of the recurrence at strip granularity.  The only
// exchange per step is one __shfl_up (the bottom row of the strip above), there is no block
// barrier, and a 128-thread CTA runs four independent problems, so a batch of small
// problems (BASELINE config 2: 41 x 150; config 5: 384 heads per utterance) fills the SMs
// with independent dependency chains.  The cost matrix is first staged into shared memory
// with coalesced 16-byte loads whenever it fits next to the trace, so no global load sits on the
// step-to-step dependency chain; larger problems (LibriSpeech-shaped, 401 x 1500) prefetch each
// lane's strip of a column kPrefetch steps ahead with cp.async into a small per-lane ring in shared
// memory (a column step is ~250 cycles, an L2 round trip ~1400: one column of lookahead left the
// sweep latency-bound).  The trace is 2 bits per cell, packed per (column, lane) word and kept in shared
// memory (global workspace only when a problem does not fit); lane 0 backtraces on the device
// and only N jump frames / W word times leave the SM.
//
// Bit-exactness contract (must match dtw_cpu): fp32 round-to-nearest add, no FMA; the
// diagonal wins only if strictly smaller than both others, then the text step only if
// strictly smaller than both others, otherwise the time step -- so ties and NaN go to
// code 2 (even when that is not the minimum).  Border rule of backtrace: column 0 -> code 1,
// row 0 -> code 2.
#include <cstdlib>

#include "common.cuh"

namespace wca {


template <int R> struct TraceWord { using type = uint64_t; };
template <> struct TraceWord<1> { using type = uint8_t; };
template <> struct TraceWord<2> { using type = uint8_t; };
template <> struct TraceWord<4> { using type = uint8_t; };
template <> struct TraceWord<8> { using type = uint16_t; };
template <> struct TraceWord<16> { using type = uint32_t; };

struct DtwLaunch {
    const float *matrix;
    const wca_utt_t *utts;
    int n_utts;
    int negate;
    int32_t *path_text, *path_time, *path_len, *jump_frames;
    const int32_t *word_bounds;
    double *start_times, *end_times;
    unsigned char *trace_ws;   // global fallback, trace_stride bytes per problem
    int64_t trace_stride;      // bytes of trace per problem (shared or global)
    int jump_stride;           // ints of jump buffer per problem in shared memory
    int trace_in_smem;
    int staged;                // 1: the whole cost matrix of a problem is staged in shared memory
    int ring;                  // 1: (not staged) each lane prefetches its strip through a cp.async ring in shared memory
    int xs_floats;             // floats of the staged matrix / the ring per problem in shared memory
    int warps;                 // problems per CTA
};

constexpr int kPrefetch = 8;  // ring slots = columns in flight per lane

__device__ __forceinline__ void cp_async4(void *smem_dst, const void *gmem_src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int kPending>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(kPending) : "memory"); }

// kTraceSmem / kStaged are compile-time so that every trace and staged-matrix access is a plain shared-memory
// instruction: a pointer selected at run time between shared and global memory is a generic pointer, and generic
// LD / ST hold their registers for hundreds of cycles (ncu source page, short-scoreboard stalls).
template <int R, bool kTraceSmem, bool kStaged>
__global__ void __launch_bounds__(128) dtw_align_kernel(const DtwLaunch p) {
    using Word = typename TraceWord<R>::type;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int prob = blockIdx.x * p.warps + warp;
    if (prob >= p.n_utts) return;  // whole warp
    const wca_utt_t u = p.utts[prob];
    const int N = u.row_end - u.row_begin;
    const int M = u.n_frames;

    // shared memory: [jump buffers][traces][staged matrices], one slice per warp
    int32_t *jump_s = reinterpret_cast<int32_t *>(smem_raw) + warp * p.jump_stride;
    unsigned char *after_jump = smem_raw + (size_t)p.warps * p.jump_stride * 4;
    Word *trace;
    if constexpr (kTraceSmem) trace = reinterpret_cast<Word *>(after_jump + (size_t)warp * p.trace_stride);
    else trace = reinterpret_cast<Word *>(p.trace_ws + (size_t)prob * p.trace_stride);
    float *xs = reinterpret_cast<float *>(after_jump + (kTraceSmem ? (size_t)p.warps * p.trace_stride : 0)) +
                (size_t)warp * p.xs_floats;
    if (N <= 0 || M <= 0) {
        if (lane == 0 && p.path_len) p.path_len[prob] = 0;
        return;
    }
    for (int r = lane; r < N; r += 32) jump_s[r] = -1;

    // ---- stage the cost matrix (coalesced, sign applied once) -----------------------------------
    const float *xg = p.matrix + u.matrix_off;
    constexpr bool staged = kStaged;
    if (staged) {
        // same misalignment in shared memory as in global memory, so the body moves 16 bytes per lane and load
        const int total = N * M;
        const int mis = (int)(((uintptr_t)xg >> 2) & 3);
        xs += mis;
        const int head = min(total, (4 - mis) & 3);
        const int body = (total - head) >> 2;
        if (lane < head) xs[lane] = p.negate ? -ld_stream(xg + lane) : ld_stream(xg + lane);
        const float4 *g4 = reinterpret_cast<const float4 *>(xg + head);
        float4 *s4 = reinterpret_cast<float4 *>(xs + head);
#pragma unroll 16
        for (int e = lane; e < body; e += 32) {  // independent loads: 8 KB of requests in flight per warp
            float4 v = ld_stream4(g4 + e);
            if (p.negate) v = make_float4(-v.x, -v.y, -v.z, -v.w);
            s4[e] = v;
        }
        for (int e = head + 4 * body + lane; e < total; e += 32) xs[e] = p.negate ? -ld_stream(xg + e) : ld_stream(xg + e);
        __syncwarp();
    }
    if constexpr (kStaged && kTraceSmem) {
        // ---- forward sweep, everything on chip: a branch-free step --------------------------------------------
        // Lane l handles table rows l*R+1 .. l*R+R at column index c = s - l (0-based).  No cell is ever skipped:
        // a lane that has not started keeps +inf in its strip (inf + 0 = inf: its x stays 0 until the first
        // predicated load), a lane past the last column or on rows >= N computes values nobody reads (rows only
        // feed the rows below them), so only the loads and the trace store are predicated and the step has no
        // divergent branch, no reconvergence barrier and no index arithmetic beyond one pointer bump.
        const int row0 = lane * R;
        float left[R], x[R];
        const float *xrow[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            left[r] = INFINITY;
            xrow[r] = xs + min(row0 + r, N - 1) * M - lane;  // xrow[r][s] = x[row][column s - lane]
            x[r] = lane == 0 ? xrow[r][0] : 0.f;             // column 0 is consumed at step 0 by lane 0 only
        }
        float diag_top = (lane == 0) ? 0.f : INFINITY;  // cost[row above strip][c]; cost[0][0] = 0 for lane 0
        Word *tcol = trace + lane - 32 * lane;          // tcol[32 * s] = trace[(s - lane) * 32 + lane]
        const int n_steps = M + (N + R - 1) / R - 1;    // the last strip with rows < N finishes column M - 1
        int c = -lane;
#pragma unroll 4
        for (int s = 0; s < n_steps; ++s, ++c) {
            float up_top = __shfl_up_sync(0xffffffffu, left[R - 1], 1);  // cost[row0][c + 1] of the table, from the strip above
            if (lane == 0) up_top = INFINITY;
            const bool act = (unsigned)c < (unsigned)M;
            const bool act_next = (unsigned)(c + 1) < (unsigned)M;
            float xc[R];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                xc[r] = x[r];
                if (act_next) x[r] = xrow[r][s + 1];  // next column in flight while this one is computed
            }
            Word tw = 0;
            float c0 = diag_top;  // cost[i-1][j-1]
            float c1 = up_top;    // cost[i-1][j]
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const float c2 = left[r];  // cost[i][j-1]
                const bool d0 = (c0 < c1) & (c0 < c2);
                const bool d1 = (c1 < c0) & (c1 < c2);
                const float cm = d0 ? c0 : (d1 ? c1 : c2);
                tw |= d0 ? (Word)0 : (d1 ? (Word)((Word)1 << (2 * r)) : (Word)((Word)2 << (2 * r)));
                const float cost = __fadd_rn(xc[r], cm);
                c0 = c2;    // this row's old value is the next row's diagonal
                c1 = cost;  // this row's new value is the next row's "up"
                left[r] = cost;
            }
            diag_top = up_top;
            if (act) tcol[32 * s] = tw;
        }
    } else {
    const bool flip = p.negate && !staged;

    // ---- forward sweep: lane l handles table rows l*R+1 .. l*R+R, column j = step - l + 1 ------
    const int row0 = lane * R;  // first 0-based matrix row of the strip
    float left[R];              // cost[i][j-1] of the strip rows, starts at cost[i][0] = inf
#pragma unroll
    for (int r = 0; r < R; ++r) left[r] = INFINITY;
    float diag_top = (lane == 0) ? 0.f : INFINITY;  // cost[row above strip][j-1]; cost[0][0] = 0 for lane 0
    float xn[R];
    // Ring mode: slot (step % kPrefetch) of this lane holds x[strip rows][column consumed at that step]
    const bool ring_mode = !staged && p.ring;
    float *ring = xs;  // [kPrefetch][R][32 lanes] floats (the staged-matrix slice is unused in this mode)
    auto prefetch_col = [&](int step) {  // column this lane consumes at `step`: j = step - lane + 1
        const int j = step - lane + 1;
        if (j >= 1 && j <= M) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int row = row0 + r;
                if (row < N) cp_async4(&ring[((step % kPrefetch) * R + r) * 32 + lane], xg + (int64_t)row * M + (j - 1));
            }
        }
        cp_async_commit();
    };
    auto load_col = [&](int j, float (&dst)[R]) {  // x[row][j-1] for the strip, 0 outside (sign applied at use)
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int row = row0 + r;
            float v = 0.f;
            if (row < N && j >= 1 && j <= M) {
                if constexpr (kStaged) v = xs[row * M + (j - 1)];
                else v = xg[(int64_t)row * M + (j - 1)];
            }
            dst[r] = v;
        }
    };
    auto take_col = [&](int step, float (&dst)[R]) {  // the ring slot of `step`, once its group has landed
        const int j = step - lane + 1;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            float v = 0.f;
            if (row0 + r < N && j >= 1 && j <= M) v = ring[((step % kPrefetch) * R + r) * 32 + lane];
            dst[r] = v;
        }
    };
    if (ring_mode) {
        for (int t = 0; t < kPrefetch; ++t) prefetch_col(t);
        cp_async_wait<kPrefetch - 1>();  // group of step 0 has landed
        take_col(0, xn);
    } else {
        load_col(1 - lane, xn);
    }
    const int n_steps = M + 31;
    for (int s = 0; s < n_steps; ++s) {
        const int j = s - lane + 1;
        // bottom row of the strip above, as of the previous step (= its column j): cost[row0][j]
        float up_top = __shfl_up_sync(0xffffffffu, left[R - 1], 1);
        if (lane == 0) up_top = INFINITY;  // cost[0][j], j >= 1
        float xc[R];
#pragma unroll
        for (int r = 0; r < R; ++r) xc[r] = flip ? -xn[r] : xn[r];
        if (ring_mode) {
            // slot of step s has been consumed into xc: refill it for step s + kPrefetch, then wait for step s + 1
            prefetch_col(s + kPrefetch);
            cp_async_wait<kPrefetch - 1>();
            take_col(s + 1, xn);
        } else {
            load_col(j + 1, xn);  // next column in flight while this one is computed
        }
        if (j >= 1 && j <= M) {
            Word tw = 0;
            float c0 = diag_top;  // cost[i-1][j-1]
            float c1 = up_top;    // cost[i-1][j]
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const float c2 = left[r];  // cost[i][j-1]
                // selects, not branches: the lanes of a warp take different cases in every cell
                const bool d0 = (c0 < c1) & (c0 < c2);
                const bool d1 = (c1 < c0) & (c1 < c2);
                const float c = d0 ? c0 : (d1 ? c1 : c2);
                const uint32_t code = d0 ? 0u : (d1 ? 1u : 2u);
                const float cost = __fadd_rn(xc[r], c);
                tw |= (Word)((Word)code << (2 * (r % (4 * (int)sizeof(Word)))));
                c0 = c2;    // this row's old value is the next row's diagonal
                c1 = cost;  // this row's new value is the next row's "up"
                left[r] = cost;
            }
            diag_top = up_top;
            trace[(int64_t)(j - 1) * 32 + lane] = tw;
        }
    }
    }
    __syncwarp();

    // ---- backtrace -------------------------------------------------------------------------------
    // Only the jump frames are wanted on the product path, and the walk leaves a text row exactly once: the whole
    // warp looks at the 32 trace cells to the left of the current one in that row, a ballot finds the nearest cell
    // whose step changes the row (code 0 / 1), and the walk moves up from there.  One shared-memory read, one
    // ballot and one shuffle per ROW (plus one per 32 consecutive time steps) instead of one dependent read per
    // path point.  The borders need no trace: a row reached at column 0 keeps the -1 its jump slot was initialised
    // with (upstream's index arithmetic gives frame -1 there), row 0 only takes time steps.
    if (!p.path_text) {
        constexpr int kShift = R == 1 ? 0 : (R == 2 ? 1 : (R == 4 ? 2 : (R == 8 ? 3 : (R == 16 ? 4 : 5))));
        int row = N - 1, col = M - 1, n_diag = 0;
        while (row >= 0 && col >= 0) {
            const int cc = col - lane;
            uint32_t code = 2u;
            if (cc >= 0) {
                const Word w = trace[cc * 32 + (row >> kShift)];
                code = (uint32_t)(w >> (2 * (row & (R - 1)))) & 3u;
            }
            const uint32_t leaves = __ballot_sync(0xffffffffu, code != 2u);
            if (leaves == 0u) {  // 32 time steps (or the left border)
                col -= 32;
                continue;
            }
            const int k = __ffs(leaves) - 1;
            const uint32_t ck = __shfl_sync(0xffffffffu, code, k);
            if (lane == 0) jump_s[row] = col - k;  // first path point of the text row
            n_diag += ck == 0u;
            col -= k + (ck == 0u);
            --row;
        }
        if (lane == 0 && p.path_len) p.path_len[prob] = N + M - n_diag;  // every diagonal step saves one path point
    } else if (lane == 0) {
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
            else {
                const int row = bi - 1;
                const Word w = trace[(int64_t)(bj - 1) * 32 + row / R];
                code = (uint32_t)(w >> (2 * (row % R))) & 3u;
            }
            // first path point of a text row: the step out of it changes the row (or ends the walk)
            if (code != 2u && bi >= 1) jump_s[bi - 1] = bj - 1;
            if (code == 0u) {
                --bi;
                --bj;
            } else if (code == 1u) {
                --bi;
            } else {
                --bj;
            }
        }
        if (p.path_len) p.path_len[prob] = cap - pos;
    }
    __syncwarp();

    // ---- jump frames and word times (timing.py:110-113) ----------------------------------------
    if (p.jump_frames)
        for (int r = lane; r < N; r += 32) p.jump_frames[u.jump_off + r] = jump_s[r];
    if (p.word_bounds && p.start_times && p.end_times) {
        const int32_t *wb = p.word_bounds + u.word_off;
        const double nan = __longlong_as_double(0x7ff8000000000000ll);
        for (int w = lane; w < u.n_words; w += 32) {
            const int a = wb[w], b = wb[w + 1];
            p.start_times[u.word_off + w] = (a >= 0 && a < N) ? (double)jump_s[a] / WCA_TOKENS_PER_SECOND : nan;
            p.end_times[u.word_off + w] = (b >= 0 && b < N) ? (double)jump_s[b] / WCA_TOKENS_PER_SECOND : nan;
        }
    }
}

// ---- several warps per problem (long texts: LibriSpeech-shaped 401 x 1500) -------------------------
// One CTA = one problem, WPP warps, global lane g = warp * 32 + lane owns R consecutive rows.  A single
// warp with R = 16 is bound by its own dependency chain (16 cells per step, ~100 cycles each); with
// 128 lanes of R = 4 the step is 4 cells and the sweep has M + 127 steps.  Inside a warp the strip above
// arrives by __shfl_up as before; between warps lane 31 of warp w publishes the bottom row of its strip,
// column by column, in a shared line `edge[w][j]` followed by a progress counter, and lane 0 of warp w + 1
// (which runs 32 steps behind by construction) polls the counter: no block barrier on the recurrence.
template <int R, int WPP, int kAhead, bool kTraceSmem>
__global__ void __launch_bounds__(32 * WPP) dtw_align_multi_kernel(const DtwLaunch p) {
    using Word = typename TraceWord<R>::type;
    constexpr int L = 32 * WPP;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = threadIdx.x;
    const int prob = blockIdx.x;
    const wca_utt_t u = p.utts[prob];
    const int N = u.row_end - u.row_begin;
    const int M = u.n_frames;
    // shared memory: [64 bytes spare][jump buffer][edge lines][trace]
    int32_t *jump_s = reinterpret_cast<int32_t *>(smem_raw + 64);
    // [WPP - 1][edge_stride] of {cost bits, column tag}: value and flag travel in ONE 8-byte store, so the
    // producer needs no fence and the consumer polls the very word it wants (tag 0 = not yet written)
    unsigned long long *edge = reinterpret_cast<unsigned long long *>(smem_raw + 64 + (size_t)p.jump_stride * 4);
    const int edge_stride = (int)(p.xs_floats / 2 / (WPP > 1 ? WPP - 1 : 1));
    Word *trace;
    if constexpr (kTraceSmem) trace = reinterpret_cast<Word *>(smem_raw + 64 + (size_t)p.jump_stride * 4 + (size_t)p.xs_floats * 4);
    else trace = reinterpret_cast<Word *>(p.trace_ws + (size_t)prob * p.trace_stride);
    if (N <= 0 || M <= 0) {
        if (g == 0 && p.path_len) p.path_len[prob] = 0;
        return;
    }
    for (int r = g; r < N; r += L) jump_s[r] = -1;
    for (int e = g; e < (WPP - 1) * edge_stride; e += L) edge[e] = 0ull;
    __syncthreads();

    const bool flip = p.negate != 0;
    const int row0 = g * R;
    float left[R];
#pragma unroll
    for (int r = 0; r < R; ++r) left[r] = INFINITY;
    float diag_top = (g == 0) ? 0.f : INFINITY;
    const float *xg = p.matrix + u.matrix_off;
    // Branch-free step (see dtw_align_kernel): rows past N read the last row (their cells feed nobody), columns
    // outside [0, M) are never loaded (the lookahead registers start at 0, which keeps a strip that has not started
    // at +inf), and only the loads, the trace store and the hand-over store are predicated.
    const float *xrow[R];  // xrow[r][s] = x[row][column s - lane] at local step s
#pragma unroll
    for (int r = 0; r < R; ++r) xrow[r] = xg + (int64_t)min(row0 + r, N - 1) * M - lane;
    // kAhead columns of lookahead in registers (ncu: 40 % of the stall samples were long-scoreboard waits on the
    // cost loads); the step loop is unrolled by kAhead so that the ring index is a compile-time constant.
    float xq[kAhead][R];
#pragma unroll
    for (int d = 0; d < kAhead; ++d)
#pragma unroll
        for (int r = 0; r < R; ++r) xq[d][r] = (unsigned)(d - lane) < (unsigned)M ? xrow[r][d] : 0.f;
    const int n_steps = M + 31;  // local steps of this warp; column of a lane: j = s - lane + 1
    const volatile unsigned long long *edge_in = edge + (size_t)(warp > 0 ? warp - 1 : 0) * edge_stride;
    volatile unsigned long long *edge_out = edge + (size_t)(warp < WPP - 1 ? warp : 0) * edge_stride;
    for (int s0 = 0; s0 < n_steps; s0 += kAhead) {
      if (warp > 0 && lane == 0 && s0 + 1 <= M) {
          // Once per group of kAhead steps: wait until the warp above has published the last column this group
          // needs, with a back-off.  Polling every step from three spinning warps kept the shared-memory pipe busy
          // and the leading warp (which every other warp waits for) spent a third of its time behind those loads.
          const int need = min(M, s0 + kAhead);
          uint32_t spins = 0;
          while ((uint32_t)(edge_in[need - 1] >> 32) != (uint32_t)need) {
              __nanosleep(40);
              if (++spins > (1u << 24)) __trap();  // a protocol bug must surface as a launch failure, not a hung GPU
          }
      }
      __syncwarp();
#pragma unroll
      for (int d = 0; d < kAhead; ++d) {
        const int s = s0 + d;
        const int c = s - lane;  // 0-based column of this lane at this step
        const bool act = (unsigned)c < (unsigned)M;
        float up_top = __shfl_up_sync(0xffffffffu, left[R - 1], 1);
        if (lane == 0) {
            up_top = INFINITY;  // cost[0][j], j >= 1 (first warp)
            if (warp > 0 && act) {
                // every word carries its own column tag: the group poll above saw the LAST column of the group, and
                // stores to different addresses need not become visible in program order, so the tag of THIS word is
                // checked too (it is almost always already there: no back-off)
                unsigned long long w = edge_in[c];
                for (uint32_t spins = 0; (uint32_t)(w >> 32) != (uint32_t)(c + 1); ++spins) {
                    if (spins > (1u << 24)) __trap();
                    w = edge_in[c];
                }
                up_top = __uint_as_float((uint32_t)w);
            }
        }
        float xc[R];
#pragma unroll
        for (int r = 0; r < R; ++r) xc[r] = flip ? -xq[d][r] : xq[d][r];
        if ((unsigned)(c + kAhead) < (unsigned)M) {
#pragma unroll
            for (int r = 0; r < R; ++r) xq[d][r] = xrow[r][s + kAhead];  // L1-cached: 8 steps share a sector
        }
        Word tw = 0;
        float c0 = diag_top;
        float c1 = up_top;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const float c2 = left[r];
            // selects, not branches: the lanes of a warp take different cases in every cell
            const bool d0 = (c0 < c1) & (c0 < c2);
            const bool d1 = (c1 < c0) & (c1 < c2);
            const float cm = d0 ? c0 : (d1 ? c1 : c2);
            tw |= d0 ? (Word)0 : (d1 ? (Word)((Word)1 << (2 * r)) : (Word)((Word)2 << (2 * r)));
            const float cost = __fadd_rn(xc[r], cm);
            c0 = c2;
            c1 = cost;
            left[r] = cost;
        }
        diag_top = up_top;
        if (act) {
            trace[c * L + g] = tw;
            if (lane == 31 && warp < WPP - 1)
                edge_out[c] = ((unsigned long long)(uint32_t)(c + 1) << 32) | __float_as_uint(left[R - 1]);
        }
      }
    }
    __syncthreads();

    if (!p.path_text) {
        // jump frames only: warp 0 walks row by row with a ballot over 32 trace cells (see dtw_align_kernel)
        if (warp == 0) {
            constexpr int kShift = R == 2 ? 1 : 2;
            static_assert(R == 2 || R == 4, "strip height of the multi-warp kernel");
            int row = N - 1, col = M - 1, n_diag = 0;
            while (row >= 0 && col >= 0) {
                const int cc = col - lane;
                uint32_t code = 2u;
                if (cc >= 0) {
                    const Word w = trace[(int64_t)cc * L + (row >> kShift)];
                    code = (uint32_t)(w >> (2 * (row & (R - 1)))) & 3u;
                }
                const uint32_t leaves = __ballot_sync(0xffffffffu, code != 2u);
                if (leaves == 0u) {
                    col -= 32;
                    continue;
                }
                const int k = __ffs(leaves) - 1;
                const uint32_t ck = __shfl_sync(0xffffffffu, code, k);
                if (lane == 0) jump_s[row] = col - k;
                n_diag += ck == 0u;
                col -= k + (ck == 0u);
                --row;
            }
            if (lane == 0 && p.path_len) p.path_len[prob] = N + M - n_diag;
        }
    } else
    if (g == 0) {
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
            else {
                const int row = bi - 1;
                const Word w = trace[(int64_t)(bj - 1) * L + row / R];
                code = (uint32_t)(w >> (2 * (row % R))) & 3u;
            }
            if (code != 2u && bi >= 1) jump_s[bi - 1] = bj - 1;
            if (code == 0u) {
                --bi;
                --bj;
            } else if (code == 1u) {
                --bi;
            } else {
                --bj;
            }
        }
        if (p.path_len) p.path_len[prob] = cap - pos;
    }
    __syncthreads();

    if (p.jump_frames)
        for (int r = g; r < N; r += L) p.jump_frames[u.jump_off + r] = jump_s[r];
    if (p.word_bounds && p.start_times && p.end_times) {
        const int32_t *wb = p.word_bounds + u.word_off;
        const double nan = __longlong_as_double(0x7ff8000000000000ll);
        for (int w = g; w < u.n_words; w += L) {
            const int a = wb[w], b = wb[w + 1];
            p.start_times[u.word_off + w] = (a >= 0 && a < N) ? (double)jump_s[a] / WCA_TOKENS_PER_SECOND : nan;
            p.end_times[u.word_off + w] = (b >= 0 && b < N) ? (double)jump_s[b] / WCA_TOKENS_PER_SECOND : nan;
        }
    }
}

// ---- long problems: lanes own strips of 8 FRAMES and sweep DOWN the text rows --------------------------------
// The row-strip kernels above read the cost matrix against its layout (one 4-byte load per lane and text row, 32
// cache lines per warp load) and keep that load on the dependency chain; for LibriSpeech-shaped problems (401 x 1500)
// they ran at ~400 ns per wavefront step.  Here the recurrence is swept the other way round: lane g owns frames
// 8g .. 8g+7 (188 lanes = 6 warps for 1500 frames) and handles text row i at step s = i + g, so
//   * a warp reads 1 KB of ONE matrix row per step, contiguous, each lane its own 32 bytes, with cp.async straight into a
//     private ring in shared memory kLongAhead rows ahead (every lane consumes what it loaded itself: no flags, no
//     barriers, just cp.async.wait_group);
//   * the sweep has N + lanes - 1 steps (~590) instead of M + lanes - 1 (~1630), of 8 cells each (a single warp issues one
//     dependent instruction every ~3 cycles, so the step time is its instruction count: 16 frames per lane and 3 warps
//     measured 300 us for 401 x 1500, see profiles/r02_dtw_long.md);
//   * the trace word of a step holds the 8 codes of the strip as two 8-bit masks (bit k: text step, bit 8+k: time
//     step; neither: diagonal): N x lanes 16-bit words in shared memory (172 KB for the largest legal problem), walked
//     by one thread with bit scans: one dependent shared-memory read per text row plus one per 8 consecutive time steps;
//   * a cell is 8 instructions with 3 of them (compare, select, add) on the chain: min(c0, c1) is the only candidate
//     that can beat the time step, and it can only if c0 and c1 are ordered and different.
// The cell rule, the fp32 add and therefore every path are those of the other kernels (and of dtw_cpu), bit for bit.
constexpr int kStripCols = 8;
constexpr int kStripChunks = kStripCols / 4;  // 16-byte pieces of a strip
constexpr int kLongAhead = 3;               // rows in flight per lane
constexpr int kLongSlots = kLongAhead + 1;  // ring slots (a power of two): the row being read is never the one being refilled
static_assert((kLongSlots & (kLongSlots - 1)) == 0, "slot index is a mask");

__device__ __forceinline__ void cp_async16_zfill(void *smem_dst, const void *gmem_src, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)),
                 "l"(gmem_src), "r"(src_bytes)
                 : "memory");
}
__device__ __forceinline__ void cp_async4_zfill(void *smem_dst, const void *gmem_src, int src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)),
                 "l"(gmem_src), "r"(src_bytes)
                 : "memory");
}

struct LongPlan {
    int wpp;       // warps per problem (lanes = 32 * wpp >= ceil(max_frames / 16))
    size_t jump, edge, ring, trace, smem;
};
static LongPlan dtw_long_plan(int max_rows, int max_frames) {
    LongPlan l;
    const int lanes_needed = (max_frames + kStripCols - 1) / kStripCols;
    l.wpp = (lanes_needed + 31) / 32;
    const size_t lanes = 32 * (size_t)l.wpp;
    l.jump = (size_t)((max_rows + 3) & ~3) * 4;
    l.edge = ((size_t)(l.wpp - 1) * (size_t)max_rows * 8 + 15) & ~(size_t)15;
    l.ring = (size_t)kLongSlots * kStripChunks * lanes * 16 + (size_t)kStripChunks * 16;  // + one zeroed chunk
    l.trace = ((size_t)max_rows * lanes * 2 + 15) & ~(size_t)15;
    l.smem = l.jump + l.edge + l.ring + l.trace;
    return l;
}

template <int WPP, bool kFlip>
__global__ void __launch_bounds__(32 * WPP) dtw_align_long_kernel(const DtwLaunch p, int max_rows) {
    constexpr int L = 32 * WPP;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = threadIdx.x;
    const int prob = blockIdx.x;
    const wca_utt_t u = p.utts[prob];
    const int N = u.row_end - u.row_begin;
    const int M = u.n_frames;
    // shared memory: [jump frames][edge words (WPP-1) x max_rows][cost ring kLongSlots x 2 x L float4][trace N x L 16-bit words]
    int32_t *jump_s = reinterpret_cast<int32_t *>(smem_raw);
    const size_t jump_bytes = (size_t)((max_rows + 3) & ~3) * 4;
    const size_t edge_bytes = ((size_t)(WPP - 1) * max_rows * 8 + 15) & ~(size_t)15;
    unsigned long long *edge = reinterpret_cast<unsigned long long *>(smem_raw + jump_bytes);
    float4 *ring = reinterpret_cast<float4 *>(smem_raw + jump_bytes + edge_bytes);
    uint16_t *trace = reinterpret_cast<uint16_t *>(reinterpret_cast<unsigned char *>(ring) + (size_t)kLongSlots * kStripChunks * L * 16 +
                                                   (size_t)kStripChunks * 16);
    if (N <= 0 || M <= 0) {
        if (g == 0 && p.path_len) p.path_len[prob] = 0;
        return;
    }
    for (int r = g; r < N; r += L) jump_s[r] = -1;
    for (int e = g; e < (WPP - 1) * N; e += L) edge[e] = 0ull;  // tag 0 = not yet published
    __syncthreads();

    const float *xg = p.matrix + u.matrix_off;
    const int j0 = g * kStripCols;                       // first frame of the strip
    const int n_cols = max(0, min(kStripCols, M - j0));  // frames of the strip inside the matrix
    // every matrix row starts 16-byte aligned (base aligned, M % 4 == 0): 16-byte copies; else 4-byte copies
    const bool vec = (reinterpret_cast<uintptr_t>(xg) & 15) == 0 && (M & 3) == 0;
    // Ring slot of row r: r % kLongSlots; chunk q of the strip at ring[(slot * kStripChunks + q) * L + g] (conflict-free
    // 16-byte reads).  Everything that does not change from step to step is hoisted: the bytes of each chunk that lie inside
    // the matrix (the rest of a copy is zero-filled), the lane's column of the ring, running offsets instead of products.
    int chunk_bytes[kStripChunks];
#pragma unroll
    for (int q = 0; q < kStripChunks; ++q) chunk_bytes[q] = 4 * max(0, min(4, n_cols - 4 * q));
    float4 *ring_lane = ring + g;
    const float4 *zero4 = ring + (size_t)kLongSlots * kStripChunks * L;  // one zeroed chunk: what lanes outside the matrix read
    if (g < kStripChunks) ring[(size_t)kLongSlots * kStripChunks * L + g] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    auto prefetch = [&](int r, int64_t row_off) {  // row r of the matrix (at float offset row_off + j0) into its ring slot
        if ((unsigned)r < (unsigned)N && n_cols > 0) {
            const float *src = xg + row_off + j0;
            float4 *dst = ring_lane + (r & (kLongSlots - 1)) * (kStripChunks * L);
            if (vec) {
#pragma unroll
                for (int q = 0; q < kStripChunks; ++q)
                    if (chunk_bytes[q] > 0) cp_async16_zfill(dst + q * L, src + 4 * q, chunk_bytes[q]);
            } else {
#pragma unroll
                for (int k = 0; k < kStripCols; ++k)
                    if (k < n_cols) cp_async4_zfill(reinterpret_cast<float *>(dst + (k >> 2) * L) + (k & 3), src + k, 4);
                    else reinterpret_cast<float *>(dst + (k >> 2) * L)[k & 3] = 0.f;
            }
        }
        cp_async_commit();
    };

    // cost[i-1][j] of the strip's frames (table row above), +inf before the first row; a lane that has not started keeps
    // +inf (inf + 0: its x reads as 0 until its first row), a lane past the last row computes values nobody reads
    float up[kStripCols];
#pragma unroll
    for (int k = 0; k < kStripCols; ++k) up[k] = INFINITY;
    float last_new = INFINITY;                      // cost[i][j0 + 7] of the row just finished (what the right neighbour needs)
    float diag_in = (g == 0) ? 0.f : INFINITY;      // cost[i-1][j0 - 1]: table cell (0, 0) = 0 for the first strip's first row
    const volatile unsigned long long *edge_in = edge + (size_t)(warp > 0 ? warp - 1 : 0) * N;
    volatile unsigned long long *edge_out = edge + (size_t)(warp < WPP - 1 ? warp : 0) * N;
    const bool warp_live = warp * 32 * kStripCols < M;  // warps whose strips all lie past the last frame only idle along
    const bool polls = lane == 0 && warp > 0, publishes = lane == 31 && warp < WPP - 1;

    // copy group n of a lane carries its row n - g: kLongAhead groups before the loop, one more per step
#pragma unroll
    for (int a = 0; a < kLongAhead; ++a) prefetch(a - g, (int64_t)(a - g) * M);
    const int n_steps = N + L - 1;
    int i = -g;                                          // text row of this lane at this step
    int64_t ahead_off = (int64_t)(kLongAhead - g) * M;   // float offset of row i + kLongAhead
    int trace_off = -g * L + g;                          // trace[i * L + g]
    if (warp_live)
    for (int s = 0; s < n_steps; ++s, ++i, ahead_off += M, trace_off += L) {
        const bool act = (unsigned)i < (unsigned)N;
        // left neighbour's newest last cell = cost[i][j0 - 1]
        float left_in = __shfl_up_sync(0xffffffffu, last_new, 1);
        if (lane == 0) left_in = INFINITY;  // table column 0 (first warp); the next warps take it from the hand-over word
        if (polls && act) {
            unsigned long long w = edge_in[i];
            for (uint32_t spins = 0; (uint32_t)(w >> 32) != (uint32_t)(i + 1); ++spins) {
                if (spins > (1u << 26)) __trap();  // a protocol bug must surface as a launch failure, not a hung GPU
                w = edge_in[i];
            }
            left_in = __uint_as_float((uint32_t)w);
        }
        cp_async_wait<kLongAhead - 1>();  // this lane's copy of row i has landed
        float x[kStripCols];
        {
            const float4 *slot = ring_lane + (i & (kLongSlots - 1)) * (kStripChunks * L);
#pragma unroll
            for (int q = 0; q < kStripChunks; ++q) {
                const float4 v = *(act ? slot + q * L : zero4);  // one select on the address instead of four on the values
                x[4 * q] = v.x; x[4 * q + 1] = v.y; x[4 * q + 2] = v.z; x[4 * q + 3] = v.w;
            }
        }
        prefetch(i + kLongAhead, ahead_off);  // into the slot read one step ago, never the one just read
        uint32_t tw_text = 0, tw_time = 0;  // bit k: the cell of frame j0 + k takes the text step / the time step
        float c0 = diag_in;   // cost[i-1][j-1]
        float c2 = left_in;   // cost[i][j-1]
#pragma unroll
        for (int k = 0; k < kStripCols; ++k) {
            const float c1 = up[k];  // cost[i-1][j]
            // dtw_cpu's rule: the diagonal only if strictly below both others, then the text step only if strictly below
            // both others, else the time step (ties and NaN).  Only min(c0, c1) can beat c2, and only when c0 and c1 are
            // ordered and different; everything but (compare with c2, select, add) is off the chain.
            const bool ordered_ne = (c0 < c1) || (c0 > c1);
            const float a = fminf(c0, c1);
            const bool win = ordered_ne && (a < c2);
            const bool text = win && (c1 < c0);
            const float cm = win ? a : c2;
            tw_time |= win ? 0u : (1u << k);
            tw_text |= text ? (1u << k) : 0u;
            const float cost = kFlip ? __fsub_rn(cm, x[k]) : __fadd_rn(x[k], cm);  // (-x) + cm, bit for bit
            c0 = c1;       // this column's old value is the next column's diagonal
            c2 = cost;     // this column's new value is the next column's left
            up[k] = cost;
        }
        diag_in = left_in;
        last_new = c2;
        if (act) {
            if (n_cols > 0) trace[trace_off] = (uint16_t)(tw_text | (tw_time << kStripCols));
            if (publishes) edge_out[i] = ((unsigned long long)(uint32_t)(i + 1) << 32) | __float_as_uint(last_new);
        }
    }
    cp_async_wait<0>();
    __syncthreads();

    // ---- backtrace by one thread: per text row, the nearest cell at or left of the current frame whose step leaves the row
    if (g == 0) {
        if (!p.path_text) {
            int row = N - 1, col = M - 1, n_diag = 0;
            while (row >= 0 && col >= 0) {
                int wi = col / kStripCols;
                const int kc = col % kStripCols;
                uint32_t w = trace[(size_t)row * L + wi];
                uint32_t leave = ~(w >> kStripCols) & 0xFFu;            // bit k set where cell k leaves the row (not a time step)
                leave &= (2u << kc) - 1u;                               // cells at or left of the current frame
                while (leave == 0u && wi > 0) {
                    --wi;
                    w = trace[(size_t)row * L + wi];
                    leave = ~(w >> kStripCols) & 0xFFu;
                }
                if (leave == 0u) break;  // left border: the remaining rows keep -1 (upstream's index arithmetic gives frame -1)
                const int k = 31 - __clz(leave);
                const bool diag = ((w >> k) & 1u) == 0u;  // leaves the row and is not a text step
                const int at = wi * kStripCols + k;
                jump_s[row] = at;       // first path point of the text row
                n_diag += diag;
                col = at - (diag ? 1 : 0);
                --row;
            }
            if (p.path_len) p.path_len[prob] = N + M - n_diag;  // every diagonal step saves one path point
        } else {
            const int cap = N + M;
            int32_t *pt = p.path_text + u.path_off, *pj = p.path_time + u.path_off;
            int bi = N, bj = M, pos = cap;
            while (bi > 0 || bj > 0) {
                --pos;
                pt[pos] = bi - 1;
                pj[pos] = bj - 1;
                uint32_t code;
                if (bj == 0) code = 1u;
                else if (bi == 0) code = 2u;
                else {
                    const uint32_t w = trace[(size_t)(bi - 1) * L + (bj - 1) / kStripCols];
                    const int k = (bj - 1) % kStripCols;
                    code = ((w >> (kStripCols + k)) & 1u) ? 2u : ((w >> k) & 1u);
                }
                if (code != 2u && bi >= 1) jump_s[bi - 1] = bj - 1;
                if (code == 0u) {
                    --bi;
                    --bj;
                } else if (code == 1u) {
                    --bi;
                } else {
                    --bj;
                }
            }
            if (p.path_len) p.path_len[prob] = cap - pos;
        }
    }
    __syncthreads();

    if (p.jump_frames)
        for (int r = g; r < N; r += L) p.jump_frames[u.jump_off + r] = jump_s[r];
    if (p.word_bounds && p.start_times && p.end_times) {
        const int32_t *wb = p.word_bounds + u.word_off;
        const double nan = __longlong_as_double(0x7ff8000000000000ll);
        for (int w = g; w < u.n_words; w += L) {
            const int a = wb[w], b = wb[w + 1];
            p.start_times[u.word_off + w] = (a >= 0 && a < N) ? (double)jump_s[a] / WCA_TOKENS_PER_SECOND : nan;
            p.end_times[u.word_off + w] = (b >= 0 && b < N) ? (double)jump_s[b] / WCA_TOKENS_PER_SECOND : nan;
        }
    }
}

// Geometry of the multi-warp variant for a launch (0 warps: use the warp-per-problem kernel).
struct MultiPlan {
    int wpp, r;
    size_t trace, edge, smem;
    int trace_in_smem;
};
static MultiPlan dtw_multi_plan(int max_rows, int max_frames) {
    MultiPlan m = {0, 0, 0, 0, 0, 0};
    if (max_rows <= 128 || max_rows > 1024) return m;
    m.wpp = max_rows <= 512 ? 4 : 8;
    const int lanes = 32 * m.wpp;
    m.r = (max_rows + lanes - 1) / lanes <= 2 ? 2 : 4;
    const size_t word = 1;  // R <= 4: one byte per (column, lane)
    m.trace = (((size_t)max_frames * lanes * word) + 15) & ~(size_t)15;
    m.edge = (size_t)(m.wpp - 1) * (((size_t)max_frames + 3) & ~(size_t)3) * 8;
    const size_t jump = (size_t)((max_rows + 3) & ~3) * 4;
    m.smem = 64 + jump + m.edge + m.trace;
    m.trace_in_smem = 1;
    if (m.smem > 227u * 1024u) {
        m.smem = 64 + jump + m.edge;
        m.trace_in_smem = 0;
    }
    return m;
}

static int strip_rows(int max_rows) {
    int r = 1;
    while (r * 32 < max_rows) r *= 2;
    return r;  // 1, 2, 4, 8, 16, 32
}
static size_t trace_word_bytes(int r) { return r <= 4 ? 1 : (r == 8 ? 2 : (r == 16 ? 4 : 8)); }
static size_t dtw_trace_bytes(int max_rows, int max_frames) {
    const size_t b = (size_t)max_frames * 32 * trace_word_bytes(strip_rows(max_rows));
    return (b + 15) & ~(size_t)15;
}
constexpr size_t kSmemBudget = 227u * 1024u;

// Shared-memory plan for a launch: problems per CTA, whether the trace, the staged matrix or the prefetch ring fit.
// The per-lane cp.async ring (prefetch_col / take_col) measured no better than plain loads one column ahead
// (64 x (401 x 1500): 1.9 ms with it, 1.5 ms without): the sweep is not bound by load latency.  Kept behind this switch.
constexpr bool kUseRing = false;
struct DtwPlan {
    int warps, trace_in_smem, staged, ring, xs_floats;
    size_t smem;
};
static DtwPlan dtw_plan(int max_rows, int max_frames) {
    const size_t jump = (size_t)((max_rows + 3) & ~3) * 4;
    const size_t trace = dtw_trace_bytes(max_rows, max_frames);
    const size_t stage = (((size_t)max_rows * max_frames + 3 + 3) & ~(size_t)3) * 4;  // + slack for the alignment shift
    const size_t ring = (size_t)kPrefetch * strip_rows(max_rows) * 32 * 4;
    DtwPlan pl;
    for (int warps = 4; warps >= 1; warps >>= 1) {  // prefer everything on chip, then more problems per CTA
        if (warps * (jump + trace + stage) <= kSmemBudget) {
            pl = {warps, 1, 1, 0, (int)(stage / 4), warps * (jump + trace + stage)};
            return pl;
        }
    }
    for (int warps = 4; warps >= 1; warps >>= 1) {
        if (kUseRing && warps * (jump + trace + ring) <= kSmemBudget) {
            pl = {warps, 1, 0, 1, (int)(ring / 4), warps * (jump + trace + ring)};
            return pl;
        }
    }
    for (int warps = 4; warps >= 1; warps >>= 1) {
        if (warps * (jump + trace) <= kSmemBudget) {
            pl = {warps, 1, 0, 0, 0, warps * (jump + trace)};
            return pl;
        }
    }
    if (kUseRing) pl = {4, 0, 0, 1, (int)(ring / 4), 4 * (jump + ring)};  // trace in the global workspace
    else pl = {4, 0, 0, 0, 0, 4 * jump};
    return pl;
}

// The column-strip kernel takes every launch whose problems are long (more than 128 text rows) or too large to stage next to
// their trace, as long as its own trace (max_rows x lanes words) fits in shared memory -- and also the launches where the staged
// warp-per-problem kernel would leave the SMs nearly empty or idle behind its shared-memory footprint: a cost matrix that
// allows fewer than four problems per CTA (probe-shaped 96 x 300: one warp per SM against ten CTAs of two warps; 1 536 problems
// 693 -> 171 us), or hundreds of problems in flight (1 680 x (41 x 150): 60.8 -> 50.8 us, 6 144: 245 -> 215 us).  A handful of
// small problems stays with the staged kernel, whose single dependency chain is shorter (16 x (41 x 150): 24 us against 28).
static bool use_long_kernel(int n_utts, int max_rows, int max_frames) {
    const LongPlan lp = dtw_long_plan(max_rows, max_frames);
    if (lp.wpp > 8 || lp.smem > kSmemBudget) return false;
    static const char *force = getenv("WCA_DTW_LONG");  // experiment switch: "1" always, "0" never (when the other kernels can)
    if (force && force[0] == '1') return true;
    if (force && force[0] == '0' && max_rows <= 1024) return false;
    if (max_rows > 128) return true;
    const DtwPlan pl = dtw_plan(max_rows, max_frames);
    return !pl.staged || pl.warps < 4 || n_utts >= 512;
}

int64_t dtw_workspace_bytes(int n_utts, int max_rows, int max_frames) {
    if (use_long_kernel(n_utts, max_rows, max_frames)) return 0;
    const MultiPlan m = dtw_multi_plan(max_rows, max_frames);
    if (m.wpp && n_utts <= 48) return m.trace_in_smem ? 0 : (int64_t)n_utts * (int64_t)m.trace;  // same rule as the launch
    if (dtw_plan(max_rows, max_frames).trace_in_smem) return 0;
    return (int64_t)n_utts * (int64_t)dtw_trace_bytes(max_rows, max_frames);
}

template <int R, int WPP, int kAhead, bool kTraceSmem>
static int launch_multi_s(const DtwLaunch &p, size_t smem, cudaStream_t stream) {
    if (smem > 48u * 1024u)
        WCA_CUDA(cudaFuncSetAttribute(dtw_align_multi_kernel<R, WPP, kAhead, kTraceSmem>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dtw_align_multi_kernel<R, WPP, kAhead, kTraceSmem><<<p.n_utts, 32 * WPP, smem, stream>>>(p);
    WCA_LAUNCH_CHECK("dtw_align_multi_kernel");
    return WCA_OK;
}
template <int R, int WPP, int kAhead>
static int launch_multi_a(const DtwLaunch &p, size_t smem, cudaStream_t stream) {
    return p.trace_in_smem ? launch_multi_s<R, WPP, kAhead, true>(p, smem, stream)
                           : launch_multi_s<R, WPP, kAhead, false>(p, smem, stream);
}
template <int R, int WPP>
static int launch_multi(const DtwLaunch &p, size_t smem, cudaStream_t stream) {
    return launch_multi_a<R, WPP, 4>(p, smem, stream);  // 4 columns of lookahead (1..8 measured the same)
}

template <int WPP, bool kFlip>
static int launch_long_f(const DtwLaunch &p, int max_rows, size_t smem, cudaStream_t stream) {
    if (smem > 48u * 1024u)
        WCA_CUDA(cudaFuncSetAttribute(dtw_align_long_kernel<WPP, kFlip>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dtw_align_long_kernel<WPP, kFlip><<<p.n_utts, 32 * WPP, smem, stream>>>(p, max_rows);
    WCA_LAUNCH_CHECK("dtw_align_long_kernel");
    return WCA_OK;
}
template <int WPP>
static int launch_long(const DtwLaunch &p, int max_rows, size_t smem, cudaStream_t stream) {
    return p.negate ? launch_long_f<WPP, true>(p, max_rows, smem, stream) : launch_long_f<WPP, false>(p, max_rows, smem, stream);
}

template <int R, bool kTraceSmem, bool kStaged>
static int launch_rs(const DtwLaunch &p, size_t smem, cudaStream_t stream) {
    if (smem > 48u * 1024u)
        WCA_CUDA(cudaFuncSetAttribute(dtw_align_kernel<R, kTraceSmem, kStaged>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)smem));
    const int grid = (p.n_utts + p.warps - 1) / p.warps;
    dtw_align_kernel<R, kTraceSmem, kStaged><<<grid, p.warps * 32, smem, stream>>>(p);
    WCA_LAUNCH_CHECK("dtw_align_kernel");
    return WCA_OK;
}
template <int R>
static int launch_r(const DtwLaunch &p, size_t smem, cudaStream_t stream) {
    if (p.trace_in_smem && p.staged) return launch_rs<R, true, true>(p, smem, stream);
    if (p.trace_in_smem) return launch_rs<R, true, false>(p, smem, stream);
    return launch_rs<R, false, false>(p, smem, stream);
}

int launch_dtw_align(const float *d_matrix, const wca_utt_t *d_utts, int n_utts, int max_rows, int max_frames,
                     int negate, int32_t *d_path_text, int32_t *d_path_time, int32_t *d_path_len,
                     int32_t *d_jump_frames, const int32_t *d_word_bounds, double *d_start_times,
                     double *d_end_times, void *d_trace_ws, int64_t trace_ws_bytes, cudaStream_t stream) {
    if (max_rows > 1024) {
        set_error("wca_dtw_align: %d text rows exceed the 1024 rows one warp sweeps (Whisper caps T at 448)", max_rows);
        return WCA_ERR_UNSUPPORTED;
    }
    if (max_rows < 1) max_rows = 1;
    if (max_frames < 1) max_frames = 1;
    DtwLaunch p;
    p.matrix = d_matrix;
    p.utts = d_utts;
    p.n_utts = n_utts;
    p.negate = negate;
    p.path_text = d_path_text;
    p.path_time = d_path_time;
    p.path_len = d_path_len;
    p.jump_frames = d_jump_frames;
    p.word_bounds = d_word_bounds;
    p.start_times = d_start_times;
    p.end_times = d_end_times;
    p.trace_ws = nullptr;
    p.jump_stride = (max_rows + 3) & ~3;
    if (use_long_kernel(n_utts, max_rows, max_frames)) {
        const LongPlan lp = dtw_long_plan(max_rows, max_frames);
        p.trace_stride = (int64_t)lp.trace;
        p.trace_in_smem = 1;
        p.staged = p.ring = 0;
        p.xs_floats = 0;
        p.warps = 1;
        switch (lp.wpp) {
            case 1: return launch_long<1>(p, max_rows, lp.smem, stream);
            case 2: return launch_long<2>(p, max_rows, lp.smem, stream);
            case 3: return launch_long<3>(p, max_rows, lp.smem, stream);
            case 4: return launch_long<4>(p, max_rows, lp.smem, stream);
            case 5: return launch_long<5>(p, max_rows, lp.smem, stream);
            case 6: return launch_long<6>(p, max_rows, lp.smem, stream);
            case 7: return launch_long<7>(p, max_rows, lp.smem, stream);
            default: return launch_long<8>(p, max_rows, lp.smem, stream);
        }
    }
    const MultiPlan mp = dtw_multi_plan(max_rows, max_frames);
    // long texts, few problems: several warps per problem shorten the dependency chain of each one; with many
    // problems in flight the warp-per-problem kernel has the higher throughput (64 x (401 x 1500): 1.48 vs 1.69 ms)
    if (mp.wpp && n_utts <= 48) {
        p.trace_stride = (int64_t)mp.trace;
        p.trace_in_smem = mp.trace_in_smem;
        p.staged = 0;
        p.ring = 0;
        p.xs_floats = (int)(mp.edge / 4);
        p.warps = 1;
        if (!p.trace_in_smem) {
            const int64_t need = (int64_t)n_utts * p.trace_stride;
            if (!d_trace_ws || trace_ws_bytes < need) {
                set_error("wca_dtw_align: trace workspace of %lld bytes required, %lld given", (long long)need,
                          (long long)trace_ws_bytes);
                return WCA_ERR_INVALID;
            }
            p.trace_ws = static_cast<unsigned char *>(d_trace_ws);
        }
        if (mp.wpp == 4) return mp.r == 2 ? launch_multi<2, 4>(p, mp.smem, stream) : launch_multi<4, 4>(p, mp.smem, stream);
        return mp.r == 2 ? launch_multi<2, 8>(p, mp.smem, stream) : launch_multi<4, 8>(p, mp.smem, stream);
    }
    p.trace_stride = (int64_t)dtw_trace_bytes(max_rows, max_frames);
    const DtwPlan pl = dtw_plan(max_rows, max_frames);
    p.trace_in_smem = pl.trace_in_smem;
    p.staged = pl.staged;
    p.ring = pl.ring;
    p.xs_floats = pl.xs_floats;
    p.warps = pl.warps;
    const size_t smem = pl.smem;
    if (!p.trace_in_smem) {
        const int64_t need = (int64_t)n_utts * p.trace_stride;
        if (!d_trace_ws || trace_ws_bytes < need) {
            set_error("wca_dtw_align: trace workspace of %lld bytes required, %lld given", (long long)need,
                      (long long)trace_ws_bytes);
            return WCA_ERR_INVALID;
        }
        p.trace_ws = static_cast<unsigned char *>(d_trace_ws);
    }
    switch (strip_rows(max_rows)) {
        case 1: return launch_r<1>(p, smem, stream);
        case 2: return launch_r<2>(p, smem, stream);
        case 4: return launch_r<4>(p, smem, stream);
        case 8: return launch_r<8>(p, smem, stream);
        case 16: return launch_r<16>(p, smem, stream);
        default: return launch_r<32>(p, smem, stream);
    }
}

}  // namespace wca
