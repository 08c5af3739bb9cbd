/* ORACLE / TEST INFRASTRUCTURE ONLY -- never linked into the shipped library.
 *
 * Plain-C restatement of the CPU dynamic-time-warping routine of the third-party
 * `openai-whisper` package (`whisper/timing.py`: dtw_cpu + backtrace, un-vendored,
 * un-pinned; the reference reaches it at timing.py:103 `dtw(-matrix)` with a CPU
 * tensor, and at timing.py:165).  Published algorithm:
 *
 *   cost  : (N+1) x (M+1) float32, +inf except cost[0][0] = 0
 *   order : column-major sweep, j = 1..M outer, i = 1..N inner
 *   pick  : c0 = cost[i-1][j-1], c1 = cost[i-1][j], c2 = cost[i][j-1]
 *           code 0 if c0 <  c1 and c0 <  c2      (diagonal)
 *           code 1 if c1 <  c0 and c1 <  c2      (text step)
 *           code 2 otherwise                     (time step; ties and NaN land here)
 *   cost[i][j] = (float)((double)x[i-1][j-1] + (double)c)   -- input is viewed as
 *           float64, the table is float32 (equals one fp32 round-to-nearest add)
 *   trace : row 0 forced to 2, column 0 forced to 1 (column assignment wins at [0][0])
 *   walk  : from (N, M) until (0, 0), emitting (i-1, j-1), then reversed.
 *
 * Also provides the boundary extraction of the reference (timing.py:110-113):
 * jumps = first path point of every run of equal text index; times = frame / 50.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

/* Returns the path length (<= N+M), or -1 on allocation failure. text_idx/time_idx
 * must hold N+M entries; they are written in forward (time-increasing) order.   */
long dtw_oracle_f32(const float *x, long N, long M, int64_t *text_idx, int64_t *time_idx)
{
    const long W = M + 1;
    float *cost = (float *)malloc(sizeof(float) * (size_t)(N + 1) * (size_t)W);
    unsigned char *trace = (unsigned char *)malloc((size_t)(N + 1) * (size_t)W);
    if (!cost || !trace) { free(cost); free(trace); return -1; }
    for (long k = 0; k < (N + 1) * W; ++k) { cost[k] = INFINITY; trace[k] = 255; }
    cost[0] = 0.0f;

    for (long j = 1; j <= M; ++j) {
        for (long i = 1; i <= N; ++i) {
            const float c0 = cost[(i - 1) * W + (j - 1)];
            const float c1 = cost[(i - 1) * W + j];
            const float c2 = cost[i * W + (j - 1)];
            float c; unsigned char t;
            if (c0 < c1 && c0 < c2)      { c = c0; t = 0; }
            else if (c1 < c0 && c1 < c2) { c = c1; t = 1; }
            else                         { c = c2; t = 2; }
            cost[i * W + j] = (float)((double)x[(i - 1) * M + (j - 1)] + (double)c);
            trace[i * W + j] = t;
        }
    }
    for (long j = 0; j <= M; ++j) trace[j] = 2;
    for (long i = 0; i <= N; ++i) trace[i * W] = 1;

    long i = N, j = M, n = 0;
    while (i > 0 || j > 0) {
        text_idx[n] = i - 1;
        time_idx[n] = j - 1;
        ++n;
        const unsigned char t = trace[i * W + j];
        if (t == 0)      { --i; --j; }
        else if (t == 1) { --i; }
        else             { --j; }
    }
    for (long a = 0, b = n - 1; a < b; ++a, --b) {
        int64_t s = text_idx[a]; text_idx[a] = text_idx[b]; text_idx[b] = s;
        s = time_idx[a]; time_idx[a] = time_idx[b]; time_idx[b] = s;
    }
    free(cost); free(trace);
    return n;
}

/* timing.py:110-111: jumps[p] = (p == 0) || text[p] != text[p-1];
 * jump_frames = time[jumps].  Returns the number of jumps written. */
long jump_frames_oracle(const int64_t *text_idx, const int64_t *time_idx, long n, int64_t *jump_frames)
{
    long k = 0;
    for (long p = 0; p < n; ++p)
        if (p == 0 || text_idx[p] != text_idx[p - 1]) jump_frames[k++] = time_idx[p];
    return k;
}

/* timing.py:111-113: start = jump[wb[w]] / 50, end = jump[wb[w+1]] / 50 (float64). */
void word_times_oracle(const int64_t *jump_frames, const int64_t *word_boundaries, long n_words,
                       double tokens_per_second, double *start_times, double *end_times)
{
    for (long w = 0; w < n_words; ++w) {
        start_times[w] = (double)jump_frames[word_boundaries[w]] / tokens_per_second;
        end_times[w] = (double)jump_frames[word_boundaries[w + 1]] / tokens_per_second;
    }
}
