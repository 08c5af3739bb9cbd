// tcgen05 / TMEM / TMA cross-attention capture (north-star kernel 1), v2: persistent,
// warp-specialised, double-buffered in tensor memory.
//
// Replaces reference timing.py:50-66 + upstream qkv_attention (SDPA off): for every
// (utterance, layer, head, block of 128 tokens) it computes
//     P[t, f] = softmax_f( median_w( (q_t * s) . (k_f * s) )[:F] * qk_scale ),  s = 64^-1/4
// and writes it ONCE, already trimmed to F, in the (L, H, T, F) layout get_attentions
// returns.  The 1500-frame logits, the torch.cat copy, the reflect-padded copy and the two
// softmax passes of the reference never touch HBM.
//
// Mapping
//   tokens  <-> TMEM lanes (UMMA M = 128): thread r of an epilogue warpgroup owns token row
//               r, so the median window and both softmax reductions are private to a thread;
//   frames  <-> TMEM columns: the fp32 accumulator IS the row buffer (filter and exp results
//               are written back in place with tcgen05.st).  512 columns = two accumulators
//               of 256 (16 halo + 224 own + 16 halo), so the tensor pipe fills one while an
//               epilogue warpgroup drains the other;
//   F > 224 : thread-block cluster of 2/4/8 CTAs along frames; per-row max and sum travel
//               through distributed shared memory, signalled with remote mbarrier arrives;
//   fp32 in, fp32-grade out on the tensor pipe: operands are split v = hi + lo (hi, lo
//               rounded to tf32) and lo*hi + hi*lo + hi*hi is accumulated: 3 x 8
//               tcgen05.mma kind::tf32 per 64-frame chunk, error ~1e-6 relative;
//   TMA     : Q / K tiles arrive as 2-D tensor-map boxes (cp.async.bulk.tensor.2d, 64 rows x
//               32 floats, 128-byte swizzle, two boxes per 64-row stage) on mbarriers into a
//               4-stage staging ring; splitter warps rewrite them into the no-swizzle K-major
//               UMMA layout [k-chunk][8-row group][8 rows][16 B].  (Row-granular
//               cp.async.bulk copies measured ~75 cycles each in the TMA unit and starved the
//               pipeline: profiles/r01_capture_tc_v2_*.)
//
// Roles of the 16 warps of a CTA (one persistent CTA per SM, static tile schedule):
//   warp 0      TMA producer            warp 1      MMA issuer (one thread)
//   warps 4-7   operand splitters       warps 8-11 / 12-15   epilogue warpgroups A / B,
//                                       alternating tiles (accumulator i & 1)
#include <cuda.h>  // CUtensorMap and its enums only; the encoder is fetched from the driver at run time

#include <cstring>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace wca {
namespace tc {

constexpr int kThreads = 512;
constexpr int kRows = 128;           // UMMA M: tokens per tile
constexpr int kChunk = 64;           // frames per MMA group (UMMA N of a full chunk)
constexpr int kHalo = 16;            // halo columns either side (>= largest half-width, 15)
constexpr int kOwnCol0 = 16;         // accumulator column of a CTA's first own frame
constexpr int kAccCols = 256;        // columns per accumulator
constexpr int kMaxOwn = kAccCols - 2 * kHalo;   // 224 own frames per CTA and tile
constexpr int kTmemCols = 512;
constexpr int kRowBytes = kHeadDim * 4;          // 256
constexpr int kStageRows = 64;
constexpr int kBoxCols = 32;                       // floats per TMA box row: 128 B, the swizzle span
constexpr int kBoxBytes = kStageRows * kBoxCols * 4;     // 8192
constexpr int kStageBytes = 2 * kBoxBytes;               // 16384: [column half][64 rows][128 B swizzled]
constexpr int kStages = 4;
constexpr int kQSplitBytes = kRows * kRowBytes;         // 32768 per hi / lo
constexpr int kKSplitBytes = kChunk * kRowBytes;        // 16384 per hi / lo per buffer
constexpr uint32_t kLboQ = kRows * 16;   // bytes between consecutive 16-byte k-chunks (A operand)
constexpr uint32_t kLboK = kChunk * 16;  // same for the B operand
constexpr uint32_t kSbo = 128;           // bytes between consecutive 8-row groups
constexpr int kTilePitch = 17;           // transpose tile pitch (odd: conflict-free column writes)
constexpr int kSplitThreads = 128;
constexpr int kEpiThreads = 128;

constexpr int kOffStage = 0;
constexpr int kOffQHi = kOffStage + kStages * kStageBytes;
constexpr int kOffQLo = kOffQHi + kQSplitBytes;
constexpr int kOffKHi = kOffQLo + kQSplitBytes;
constexpr int kOffKLo = kOffKHi + 2 * kKSplitBytes;
constexpr int kOffTile = kOffKLo + 2 * kKSplitBytes;            // 8 epilogue warps x 32 x 17 floats
constexpr int kOffStat = kOffTile + 8 * 32 * kTilePitch * 4;    // smax[2][128], ssum[2][128]
constexpr int kOffBar = kOffStat + 4 * kRows * 4;
enum Bar {
    kStageFull = 0,
    kStageEmpty = kStageFull + kStages,
    kAReady = kStageEmpty + kStages,
    kAFree = kAReady + 1,
    kBReady = kAFree + 1,
    kBFree = kBReady + 2,
    kAccFull = kBFree + 2,
    kAccEmpty = kAccFull + 2,
    kXMax = kAccEmpty + 2,
    kXSum = kXMax + 2,
    kNumBars = kXSum + 2
};
constexpr int kOffTmem = kOffBar + kNumBars * 8;
constexpr int kSmemBytes = kOffTmem + 16;
static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");

// ------------------------------------------------------------------ debug trace (WCA_CAPTURE_TRACE)
// CTA 0 stamps clock64() at the hand-over points of every role for its first kTraceTiles
// tiles; tools/trace_capture.py turns the table into a per-role timeline.
constexpr int kTraceTiles = 40;
enum Ev { kEvProdQ = 0, kEvProdK0, kEvProdKLast, kEvSplAFree, kEvSplQDone, kEvSplK0Done, kEvSplKLast, kEvMmaAccEmpty,
          kEvMmaAReady, kEvMmaB0, kEvMmaIssued, kEvEpiAccFull, kEvEpiA, kEvEpiXMax, kEvEpiB, kEvEpiXSum, kEvEpiC,
          kEvSplFull, kEvSplLoaded, kEvSplStored, kEvSplFenced, kNumEv };
__device__ long long g_trace[kTraceTiles][kNumEv];
__device__ __forceinline__ void stamp(bool on, uint32_t tile_seq, int ev) {
    if (on && tile_seq < (uint32_t)kTraceTiles) g_trace[tile_seq][ev] = clock64();
}

// ------------------------------------------------------------------ tile geometry
struct TensorMaps {
    CUtensorMap q[WCA_MAX_LAYERS];  // (rows, H*64) fp32 matrices, box 64 rows x 32 floats, 128B swizzle
    CUtensorMap k[WCA_MAX_LAYERS];
};

struct KernelArgs {
    const wca_utt_t *utts;
    float *ws;
    int n_heads, lh_count, tok_blocks, n_tiles;
    float s, qk_scale;
    int raw_logits;
    int trace;
    unsigned dbg;  // experiment switches (bits 8..): never set by the product path
};

struct Geo {
    bool live;       // the tile has token rows (same answer in every CTA of the cluster)
    int T, F, rows_valid;
    int f0, f1, n_own;       // frames this CTA owns
    int half;                // filter half-width actually applied (0: identity)
    int m0, mcol0, n_mma;    // first frame / accumulator column / frame count the MMA computes
    int n_chunks;
    bool dup;                // <= 64 token rows and a single-CTA cluster: rows are mirrored into A rows 64..127
    int layer, col0;         // decoder layer and first float column of the head
    int qrow0, krow0;        // first Q row of the tile / first K row of the utterance
    float *out;              // row 0 of this tile, frame 0
};

template <int W>
__device__ __forceinline__ Geo decode_tile(const KernelArgs &a, int tile, uint32_t crank, uint32_t csize) {
    Geo g;
    const int tb = tile % a.tok_blocks;
    const int lh = (tile / a.tok_blocks) % a.lh_count;
    const int ub = tile / (a.tok_blocks * a.lh_count);
    const wca_utt_t u = a.utts[ub];
    g.T = u.n_tokens;
    g.F = u.n_frames;
    const int t0 = tb * kRows;
    g.live = t0 < g.T;
    g.rows_valid = min(kRows, g.T - t0);
    g.layer = lh / a.n_heads;
    g.col0 = (lh % a.n_heads) * kHeadDim;
    g.qrow0 = (int)u.q_row0 + t0;
    g.krow0 = (int)u.k_row0;
    g.out = a.ws + u.ws_off + ((int64_t)lh * g.T + t0) * g.F;
    const int slab = (((g.F + (int)csize - 1) / (int)csize) + 15) & ~15;
    g.f0 = (int)crank * slab;
    g.f1 = min(g.F, g.f0 + slab);
    g.n_own = max(0, g.f1 - g.f0);
    const bool filter = !a.raw_logits && W > 1 && g.F > W / 2;  // identity for very short rows, like upstream
    g.half = filter ? W / 2 : 0;
    g.m0 = g.f0 > 0 ? g.f0 - kHalo : 0;
    g.mcol0 = g.f0 > 0 ? 0 : kOwnCol0;  // frame f always sits at accumulator column f - f0 + 16
    g.n_mma = g.n_own > 0 ? min(g.f1 + kHalo, g.F) - g.m0 : 0;
    g.n_chunks = (g.n_mma + kChunk - 1) / kChunk;
    g.dup = csize == 1 && g.rows_valid <= kStageRows;
    return g;
}

// Staging items of one tile, in ring order: K chunk 0, then the Q halves, then K chunks 1..
// (K0 first so that it can be split while the previous tile's MMAs still read the Q buffers).
struct Item {
    bool is_q;
    int index;  // Q half (0/1) or K chunk
};
__device__ __forceinline__ int tile_items(const Geo &g) { return g.n_chunks + (g.rows_valid > kStageRows ? 2 : 1); }
__device__ __forceinline__ Item tile_item(const Geo &g, int i) {
    const int n_q = g.rows_valid > kStageRows ? 2 : 1;
    Item it;
    it.is_q = i >= 1 && i <= n_q;
    it.index = i == 0 ? 0 : (it.is_q ? i - 1 : i - n_q);
    return it;
}

// ------------------------------------------------------------------ role bodies
// One elected lane arms the stage barrier and issues the two column-half boxes of a 64-row tile.
__device__ __forceinline__ void producer_item(uint32_t stage, uint32_t bar_full, uint32_t bar_empty, uint32_t n_item,
                                              const CUtensorMap *map, int col0, int row0, int lane) {
    mbar_wait(bar_empty, ((n_item / kStages) & 1u) ^ 1u);  // first lap passes immediately
    if (lane == 0) {
        mbar_expect_tx(bar_full, kStageBytes);  // boxes are always full: rows past the matrix end read as zero
        tma_load_box(stage, map, col0, row0, bar_full);
        tma_load_box(stage + kBoxBytes, map, col0 + kBoxCols, row0, bar_full);
    }
    __syncwarp();
}

// 128 splitter threads: staging tile (64 rows) -> scaled hi / lo parts in UMMA no-swizzle layout.
__device__ __forceinline__ void split_stage(const unsigned char *stage, unsigned char *hi, unsigned char *lo,
                                            uint32_t lbo, int row_off, float s, int t, bool mirror = false,
                                            bool tr = false, uint32_t seq = 0) {
    constexpr int kIters = (kStageRows * 16) / kSplitThreads;  // 8 float4 per thread
    // e = it * 128 + t  ->  row = t & 63 (fixed per thread), 16-byte chunk ch = 2 * it + (t >> 6)
    const int row = t & (kStageRows - 1);
    const int ch0 = t >> 6;
    float4 v[kIters];
#pragma unroll
    for (int it = 0; it < kIters; ++it) {  // all loads in flight before any arithmetic
        const int ch = 2 * it + ch0;
        // 128-byte swizzle of the TMA box: 16-byte chunk c of row r sits at chunk position c ^ (r & 7)
        v[it] = *reinterpret_cast<const float4 *>(stage + (ch >> 3) * kBoxBytes + row * 128 + (((ch & 7) ^ (row & 7)) << 4));
    }
    if (tr && v[0].x != 1e30f) stamp(tr, seq, kEvSplLoaded);
#pragma unroll
    for (int it = 0; it < kIters; ++it) {
        const int ch = 2 * it + ch0;
        const float x[4] = {v[it].x * s, v[it].y * s, v[it].z * s, v[it].w * s};
        float h[4], l[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            // round-to-nearest onto the 10-bit tf32 mantissa; x - h is exact in fp32 (may be negative)
            h[c] = __uint_as_float((__float_as_uint(x[c]) + 0x1000u) & 0xFFFFE000u);
            l[c] = __uint_as_float((__float_as_uint(x[c] - h[c]) + 0x1000u) & 0xFFFFE000u);
        }
        const uint32_t off = (uint32_t)ch * lbo + (uint32_t)(row + row_off) * 16u;
        *reinterpret_cast<float4 *>(hi + off) = make_float4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<float4 *>(lo + off) = make_float4(l[0], l[1], l[2], l[3]);
        if (mirror) {  // second copy of the token rows in A rows 64..127 (see Geo::dup)
            *reinterpret_cast<float4 *>(hi + off + kStageRows * 16u) = make_float4(h[0], h[1], h[2], h[3]);
            *reinterpret_cast<float4 *>(lo + off + kStageRows * 16u) = make_float4(l[0], l[1], l[2], l[3]);
        }
    }
    stamp(tr, seq, kEvSplStored);
    fence_proxy_async_smem();  // generic-proxy stores must be visible to the tensor core's async proxy
    stamp(tr, seq, kEvSplFenced);
}

// Sliding median over a 16-column block with its neighbour blocks; every index is a compile
// time constant after unrolling, so the window lives in registers.
template <int W>
__device__ __forceinline__ void median_block(const float (&prev)[16], const float (&cur)[16], const float (&next)[16],
                                             float (&out)[16]) {
    constexpr int half = W / 2;
#pragma unroll
    for (int pos = 0; pos < 16; ++pos) {
        if constexpr (W == 1) {
            out[pos] = cur[pos];
        } else {
            float win[W];
#pragma unroll
            for (int j = 0; j < W; ++j) {
                const int idx = 16 + pos - half + j;  // position in prev | cur | next
                win[j] = idx < 16 ? prev[idx] : (idx < 32 ? cur[idx - 16] : next[idx - 32]);
            }
            out[pos] = median_regs<W>(win);
        }
    }
}

// The three sweeps of one tile by one epilogue warpgroup (thread <-> token row).
template <int W>
__device__ __forceinline__ void epilogue_tile(const Geo &g, const KernelArgs &a, unsigned char *smem, uint32_t acc,
                                              int grp, int ewarp, int lane, uint32_t csize, uint32_t x_parity,
                                              bool tr, uint32_t seq) {
    // Mirrored tiles (Geo::dup): lane quarters 2 and 3 hold a copy of token rows 0..63, so warps
    // 2/3 (other two schedulers) take the second half of the columns of the rows of warps 0/1.
    const int lw = g.dup ? (ewarp & 1) : ewarp;  // logical 32-row group
    const int row = lw * 32 + lane;
    const bool row_ok = row < g.rows_valid;
    const int n_blocks = (g.n_own + 15) >> 4;
    const int split = g.dup ? (n_blocks + 1) >> 1 : n_blocks;
    const int b_lo = (g.dup && ewarp >= 2) ? split : 0;
    const int b_hi = (g.dup && ewarp >= 2) ? n_blocks : split;
    const bool rows_live = (lw * 32 < g.rows_valid) && g.n_own > 0;  // warp-uniform
    const bool sweep = rows_live && b_lo < b_hi;
    const uint32_t trow = acc + ((uint32_t)(ewarp * 32) << 16) + kOwnCol0;  // own frame 0 of this lane quarter
    const int tail = g.n_own - 16 * (n_blocks - 1);  // valid columns of the last block (1..16)
    float *smax = reinterpret_cast<float *>(smem + kOffStat) + grp * kRows;
    float *ssum = reinterpret_cast<float *>(smem + kOffStat) + (2 + grp) * kRows;
    const uint32_t bar_xmax = smem_u32(smem + kOffBar) + 8u * (kXMax + grp);
    const uint32_t bar_xsum = smem_u32(smem + kOffBar) + 8u * (kXSum + grp);

    float inv_sum = 1.f;
    if (!a.raw_logits) {
        float row_max = -INFINITY;
        if (rows_live && g.half > 0) {
            {
                // materialise the reflect padding in TMEM: frame -i <- frame i, frame F-1+i <- frame F-1-i
                if (g.f0 == 0)
                    for (int i = 1; i <= g.half; ++i) tmem_st1(trow - (uint32_t)i, tmem_ld1(trow + (uint32_t)i));
                for (int f = max(g.F, g.f1); f < g.f1 + g.half; ++f)
                    tmem_st1(trow + (uint32_t)(f - g.f0), tmem_ld1(trow + (uint32_t)(2 * (g.F - 1) - f - g.f0)));
                tmem_wait_st();
            }
        }
        if (sweep) {
            // sweep A: median filter in place, * qk_scale, running max; the load of block b+2 is in flight
            float prev[16], cur[16], next[16], ahead[16], med[16];
            tmem_ld16_issue(trow + (uint32_t)(16 * (b_lo - 1) + 16) - 16u, prev);
            tmem_ld16_issue(trow + (uint32_t)(16 * b_lo), cur);
            tmem_ld16_issue(trow + (uint32_t)(16 * (b_lo + 1)), next);
            tmem_ld_wait(prev);
            tmem_ld_wait(cur);
            tmem_ld_wait(next);
#pragma unroll(W <= 3 ? 4 : 1)  // 4 = rotation period of prev/cur/next/ahead: the copies vanish
            for (int b = b_lo; b < b_hi; ++b) {
                if (b + 2 <= n_blocks) tmem_ld16_issue(trow + (uint32_t)(16 * (b + 2)), ahead);  // stays inside the accumulator
                if (g.half > 0) {
                    median_block<W>(prev, cur, next, med);
                } else {
#pragma unroll
                    for (int i = 0; i < 16; ++i) med[i] = cur[i];
                }
#pragma unroll
                for (int i = 0; i < 16; ++i) med[i] *= a.qk_scale;
                if (b + 1 < n_blocks) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) row_max = fmaxf(row_max, med[i]);
                } else {
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        if (i < tail) row_max = fmaxf(row_max, med[i]);
                }
                tmem_st16(trow + (uint32_t)(16 * b), med);
                tmem_ld_wait(ahead);
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    prev[i] = cur[i];
                    cur[i] = next[i];
                    next[i] = ahead[i];
                }
            }
            tmem_wait_st();
        }
        stamp(tr, seq, kEvEpiA);
        float gmax = row_max;
        if (g.dup) {  // combine the two column halves of a mirrored tile
            smax[ewarp * 32 + lane] = row_max;
            named_bar_sync(1 + grp, kEpiThreads);
            gmax = fmaxf(row_max, smax[(ewarp ^ 2) * 32 + lane]);
        }
        if (csize > 1) {
            // exchange the row maxima across the cluster (every CTA takes part, even with no own frames)
            smax[row] = row_ok ? row_max : -INFINITY;
            named_bar_sync(1 + grp, kEpiThreads);
            if (row == 0)
                for (uint32_t r = 0; r < csize; ++r) mbar_arrive_remote(bar_xmax, r);
            mbar_wait_cluster(bar_xmax, x_parity);
            gmax = -INFINITY;
            for (uint32_t r = 0; r < csize; ++r) gmax = fmaxf(gmax, ld_dsmem_f32(&smax[row], r));
        }
        stamp(tr, seq, kEvEpiXMax);

        float row_sum = 0.f;
        if (sweep) {
            // sweep B: e = exp(x - max) in place, running sum; one block of loads in flight
            const float kLog2e = 1.4426950408889634f;
            const float shift = gmax * kLog2e;
            float v[16], ahead[16];
            tmem_ld16_issue(trow + (uint32_t)(16 * b_lo), v);
            tmem_ld_wait(v);
#pragma unroll 2
            for (int b = b_lo; b < b_hi; ++b) {
                tmem_ld16_issue(trow + (uint32_t)(16 * (b + 1)), ahead);
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = ex2_approx(fmaf(v[i], kLog2e, -shift));
                if (b + 1 < n_blocks) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) row_sum += v[i];
                } else {
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        if (i < tail) row_sum += v[i];
                }
                tmem_st16(trow + (uint32_t)(16 * b), v);
                tmem_ld_wait(ahead);
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = ahead[i];
            }
            tmem_wait_st();
        }
        stamp(tr, seq, kEvEpiB);
        float gsum = row_sum;
        if (g.dup) {
            ssum[ewarp * 32 + lane] = row_sum;
            named_bar_sync(1 + grp, kEpiThreads);
            gsum = row_sum + ssum[(ewarp ^ 2) * 32 + lane];
        }
        if (csize > 1) {
            ssum[row] = row_ok ? row_sum : 0.f;
            named_bar_sync(1 + grp, kEpiThreads);
            if (row == 0)
                for (uint32_t r = 0; r < csize; ++r) mbar_arrive_remote(bar_xsum, r);
            mbar_wait_cluster(bar_xsum, x_parity);
            gsum = 0.f;
            for (uint32_t r = 0; r < csize; ++r) gsum += ld_dsmem_f32(&ssum[row], r);
        }
        inv_sum = 1.f / gsum;
        stamp(tr, seq, kEvEpiXSum);
    }

    // sweep C: normalise, transpose 32x16 blocks through shared memory, coalesced row stores.
    // Lane (c, rsel) stores column c of rows rsel, rsel+2, ...: two 64-byte row segments per
    // instruction, the address advancing by two rows per step.
    if (sweep) {
        float *tile = reinterpret_cast<float *>(smem + kOffTile) + (grp * 4 + ewarp) * (32 * kTilePitch);
        const int rows_here = min(32, g.rows_valid - lw * 32);
        const int c = lane & 15, rsel = lane >> 4;
        const int n_steps = (rows_here - rsel + 1) >> 1;  // rows rsel + 2k < rows_here
        const int64_t step = 2 * (int64_t)g.F;
        float *obase = g.out + (int64_t)(lw * 32 + rsel) * g.F + g.f0 + c;
        const float *tsrc = tile + rsel * kTilePitch + c;
        float v[16], ahead[16];
        tmem_ld16_issue(trow + (uint32_t)(16 * b_lo), v);
        tmem_ld_wait(v);
        for (int b = b_lo; b < b_hi; ++b) {
            tmem_ld16_issue(trow + (uint32_t)(16 * (b + 1)), ahead);
#pragma unroll
            for (int i = 0; i < 16; ++i) tile[lane * kTilePitch + i] = v[i] * inv_sum;
            __syncwarp();
            float o[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) o[k] = tsrc[k * 2 * kTilePitch];  // all reads in flight before the stores
            if (g.f0 + 16 * b + c < g.f1) {
                float *p = obase + 16 * b;
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                    if (k < n_steps) st_stream(p, o[k]);
                    p += step;
                }
            }
            __syncwarp();
            tmem_ld_wait(ahead);
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = ahead[i];
        }
    }
    stamp(tr, seq, kEvEpiC);
}

// ------------------------------------------------------------------ the kernel
template <int W>
__global__ void __launch_bounds__(kThreads, 1)
capture_tc_kernel(const __grid_constant__ TensorMaps maps, const __grid_constant__ KernelArgs a) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t crank = cluster_ctarank(), csize = cluster_nctarank();
    const int cid = blockIdx.x / csize, n_clusters = gridDim.x / csize;
    const uint32_t bars = smem_u32(smem + kOffBar);
    auto bar = [&](int which) { return bars + 8u * (uint32_t)which; };
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + kOffTmem);

    // ---- setup ------------------------------------------------------------------------
    if (tid == 0) {
        for (int i = 0; i < kStages; ++i) {
            mbar_init(bar(kStageFull + i), 1);
            mbar_init(bar(kStageEmpty + i), kSplitThreads);
        }
        mbar_init(bar(kAReady), kSplitThreads);
        mbar_init(bar(kAFree), 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(bar(kBReady + i), kSplitThreads);
            mbar_init(bar(kBFree + i), 1);
            mbar_init(bar(kAccFull + i), 1);
            mbar_init(bar(kAccEmpty + i), kEpiThreads);
            mbar_init(bar(kXMax + i), csize);
            mbar_init(bar(kXSum + i), csize);
        }
        fence_mbar_init();
    }
    if (warp == 0) {
        tmem_alloc(smem_u32(tmem_slot), kTmemCols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    cluster_sync_all();  // peers' barriers are initialised before anyone arrives remotely

    if (warp == 0) {
        // ================= TMA producer =================
        uint32_t n_item = 0, seq = 0;
        const bool tr = a.trace && blockIdx.x == 0 && lane == 0;
        const uint32_t stage0 = smem_u32(smem + kOffStage);
        for (int tile = cid; tile < a.n_tiles; tile += n_clusters) {
            const Geo g = decode_tile<W>(a, tile, crank, csize);
            if (!g.live || g.n_own == 0) continue;
            const int n_items = tile_items(g);
            for (int i = 0; i < n_items; ++i) {
                const Item item = tile_item(g, i);
                const uint32_t s = n_item % kStages;
                if (item.is_q)
                    producer_item(stage0 + s * kStageBytes, bar(kStageFull + s), bar(kStageEmpty + s), n_item,
                                  &maps.q[g.layer], g.col0, g.qrow0 + item.index * kStageRows, lane);
                else
                    producer_item(stage0 + s * kStageBytes, bar(kStageFull + s), bar(kStageEmpty + s), n_item,
                                  &maps.k[g.layer], g.col0, g.krow0 + g.m0 + item.index * kChunk, lane);
                ++n_item;
                stamp(tr, seq, item.is_q ? kEvProdQ : (item.index == 0 ? kEvProdK0 : kEvProdKLast));
            }
            ++seq;
        }
    } else if (warp == 1) {
        // ================= MMA issuer: warp-uniform control flow, one elected lane issues =================
        uint32_t n_tile = 0, n_chunk = 0, acc_use[2] = {0, 0}, it = 0;
        const bool tr = a.trace && blockIdx.x == 0 && lane == 0;
        const uint32_t a_hi = smem_u32(smem + kOffQHi), a_lo = smem_u32(smem + kOffQLo);
        for (int tile = cid; tile < a.n_tiles; tile += n_clusters) {
            const Geo g = decode_tile<W>(a, tile, crank, csize);
            if (!g.live) continue;
            const uint32_t buf = it & 1u;
            ++it;
            if (g.n_own == 0) continue;
            mbar_wait(bar(kAccEmpty + buf), (acc_use[buf] & 1u) ^ 1u);  // epilogue drained this accumulator
            ++acc_use[buf];
            stamp(tr, n_tile, kEvMmaAccEmpty);
            mbar_wait(bar(kAReady), n_tile & 1u);
            stamp(tr, n_tile, kEvMmaAReady);
            ++n_tile;
            for (int j = 0; j < g.n_chunks; ++j) {
                const uint32_t kb = n_chunk & 1u;
                mbar_wait(bar(kBReady + kb), (n_chunk >> 1) & 1u);
                ++n_chunk;
                if (j == 0) stamp(tr, n_tile - 1, kEvMmaB0);
                tc_fence_after();
                const int n_cols = (min(kChunk, g.n_mma - j * kChunk) + 15) & ~15;  // UMMA N: multiple of 16
                const uint32_t idesc = instr_desc_tf32(kRows, n_cols);
                const uint32_t d = tmem_base + buf * kAccCols + (uint32_t)(g.mcol0 + j * kChunk);
                const uint32_t b_hi = smem_u32(smem + kOffKHi) + kb * kKSplitBytes;
                const uint32_t b_lo = smem_u32(smem + kOffKLo) + kb * kKSplitBytes;
                if (elect_one()) {
                    // small terms first: lo*hi, hi*lo, then hi*hi
#pragma unroll
                    for (int pass = 0; pass < 3; ++pass) {
                        const uint64_t da = smem_desc(pass == 0 ? a_lo : a_hi, kLboQ, kSbo);
                        const uint64_t db = smem_desc(pass == 1 ? b_lo : b_hi, kLboK, kSbo);
#pragma unroll
                        for (int ks = 0; ks < kHeadDim / 8; ++ks)  // K = 8 tf32 = two 16-byte k-chunks per instruction
                            umma_tf32(d, da + (uint64_t)(ks * ((2 * kLboQ) >> 4)), db + (uint64_t)(ks * ((2 * kLboK) >> 4)),
                                      idesc, (pass | ks) != 0);
                    }
                    umma_commit(bar(kBFree + kb));
                    if (j == g.n_chunks - 1) {
                        umma_commit(bar(kAFree));
                        umma_commit(bar(kAccFull + buf));
                    }
                }
                __syncwarp();
            }
            stamp(tr, n_tile - 1, kEvMmaIssued);
        }
    } else if (warp >= 4 && warp < 8) {
        // ================= operand splitters =================
        const int t = tid - 4 * 32;
        uint32_t n_item = 0, n_tile = 0, n_chunk = 0;
        const bool tr = a.trace && blockIdx.x == 0 && t == 0;
        for (int tile = cid; tile < a.n_tiles; tile += n_clusters) {
            const Geo g = decode_tile<W>(a, tile, crank, csize);
            if (!g.live || g.n_own == 0) continue;
            const int n_items = tile_items(g);
            const int n_q = g.rows_valid > kStageRows ? 2 : 1;
            for (int i = 0; i < n_items; ++i) {
                const Item item = tile_item(g, i);
                const uint32_t s = n_item % kStages;
                if (item.is_q) {
                    if (item.index == 0) {
                        mbar_wait(bar(kAFree), (n_tile & 1u) ^ 1u);  // MMAs of the previous tile no longer read Q
                        stamp(tr, n_tile, kEvSplAFree);
                        ++n_tile;
                    }
                    mbar_wait(bar(kStageFull + s), (n_item / kStages) & 1u);
                    stamp(tr, n_tile - 1, kEvSplFull);
                    split_stage(smem + kOffStage + s * kStageBytes, smem + kOffQHi, smem + kOffQLo, kLboQ,
                                item.index * kStageRows, a.s, t, g.dup, tr, n_tile - 1);
                    mbar_arrive(bar(kStageEmpty + s));
                    if (item.index == n_q - 1) {
                        mbar_arrive(bar(kAReady));
                        stamp(tr, n_tile - 1, kEvSplQDone);
                    }
                } else {
                    const uint32_t kb = n_chunk & 1u;
                    mbar_wait(bar(kBFree + kb), ((n_chunk >> 1) & 1u) ^ 1u);  // MMA released this K buffer
                    mbar_wait(bar(kStageFull + s), (n_item / kStages) & 1u);
                    split_stage(smem + kOffStage + s * kStageBytes, smem + kOffKHi + kb * kKSplitBytes,
                                smem + kOffKLo + kb * kKSplitBytes, kLboK, 0, a.s, t);
                    mbar_arrive(bar(kStageEmpty + s));
                    mbar_arrive(bar(kBReady + kb));
                    stamp(tr, n_tile - (item.index == 0 ? 0 : 1), item.index == 0 ? kEvSplK0Done : kEvSplKLast);
                    ++n_chunk;
                }
                ++n_item;
            }
        }
    } else if (warp >= 8) {
        // ================= epilogue warpgroups (A: even tiles, B: odd tiles) =================
        const int grp = (warp - 8) >> 2, ewarp = warp & 3;
        uint32_t it = 0, acc_use = 0, n_x = 0;
        const bool tr = a.trace && blockIdx.x == 0 && ewarp == 0 && lane == 0;
        for (int tile = cid; tile < a.n_tiles; tile += n_clusters) {
            const Geo g = decode_tile<W>(a, tile, crank, csize);
            if (!g.live) continue;
            const bool mine = (it & 1u) == (uint32_t)grp;
            const uint32_t seq = it;
            ++it;
            if (!mine) continue;
            if (g.n_own > 0) {
                mbar_wait(bar(kAccFull + grp), acc_use & 1u);
                tc_fence_after();
            }
            stamp(tr, seq, kEvEpiAccFull);
            epilogue_tile<W>(g, a, smem, tmem_base + (uint32_t)grp * kAccCols, grp, ewarp, lane, csize, n_x & 1u, tr, seq);
            ++n_x;
            if (g.n_own > 0) {
                tc_fence_before();
                mbar_arrive(bar(kAccEmpty + grp));
                ++acc_use;
            }
        }
    }

    // ---- teardown: nobody leaves while a peer may still read our statistics ----------------
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

}  // namespace tc

int read_capture_trace(long long *h_out, int capacity) {
    const int n = tc::kTraceTiles * tc::kNumEv;
    if (capacity < n) {
        set_error("wca_debug_capture_trace: capacity %d < %d", capacity, n);
        return WCA_ERR_INVALID;
    }
    WCA_CUDA(cudaMemcpyFromSymbol(h_out, tc::g_trace, sizeof(long long) * n));
    return n;
}

bool capture_tc_supported(int max_tokens, int max_frames, int medfilt_width) {
    (void)max_tokens;
    const bool width_ok = medfilt_width == 1 || medfilt_width == 3 || medfilt_width == 5 || medfilt_width == 7;
    return width_ok && max_frames <= 8 * tc::kMaxOwn;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int encode_map(EncodeTiledFn encode, CUtensorMap *map, const float *base, int64_t rows, int64_t cols, int64_t ld) {
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
    const cuuint32_t box[2] = {(cuuint32_t)tc::kBoxCols, (cuuint32_t)tc::kStageRows};
    const cuuint32_t elem[2] = {1, 1};
    const CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(base), dims, strides, box, elem,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows %lld, cols %lld, ld %lld)", (int)r, (long long)rows,
                  (long long)cols, (long long)ld);
        return WCA_ERR_CUDA;
    }
    return WCA_OK;
}

int launch_capture_tc(const float *const *h_q_layers, const float *const *h_k_layers, int n_layers, int n_heads,
                      int64_t ld_q, int64_t ld_k, int64_t q_rows, int64_t k_rows, const wca_utt_t *d_utts, int n_utts,
                      int max_tokens, int max_frames, int medfilt_width, float qk_scale, float *d_ws, unsigned flags,
                      int sm_count, cudaStream_t stream) {
    static EncodeTiledFn encode = nullptr;
    if (!encode) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        WCA_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        if (qres != cudaDriverEntryPointSuccess || !fn) {
            set_error("capture_tc: the driver does not export cuTensorMapEncodeTiled");
            return WCA_ERR_CUDA;
        }
        encode = reinterpret_cast<EncodeTiledFn>(fn);
    }
    tc::TensorMaps maps;
    memset(&maps, 0, sizeof(maps));
    for (int l = 0; l < n_layers; ++l) {
        int rc = encode_map(encode, &maps.q[l], h_q_layers[l], q_rows, (int64_t)n_heads * kHeadDim, ld_q);
        if (rc) return rc;
        rc = encode_map(encode, &maps.k[l], h_k_layers[l], k_rows, (int64_t)n_heads * kHeadDim, ld_k);
        if (rc) return rc;
    }
    int csize = 1;
    while (csize < 8 && ((((max_frames + csize - 1) / csize) + 15) & ~15) > tc::kMaxOwn) csize *= 2;
    const int tok_blocks = (max_tokens + tc::kRows - 1) / tc::kRows;
    const int lh_count = n_layers * n_heads;
    const long long tiles = (long long)n_utts * lh_count * tok_blocks;
    if (tiles > 0x7fffffffLL) {
        set_error("capture_tc: %lld tiles exceed the tile index range", tiles);
        return WCA_ERR_UNSUPPORTED;
    }
    tc::KernelArgs a;
    a.utts = d_utts;
    a.ws = d_ws;
    a.n_heads = n_heads;
    a.lh_count = lh_count;
    a.tok_blocks = tok_blocks;
    a.n_tiles = (int)tiles;
    a.s = (float)0.35355339059327373;  // 64 ** -0.25 as the reference's fp32 scalar
    a.qk_scale = qk_scale;
    a.raw_logits = (flags & WCA_CAPTURE_RAW_LOGITS) ? 1 : 0;
    a.trace = (flags & WCA_CAPTURE_TRACE) ? 1 : 0;
    a.dbg = flags & 0xff00u;
    const int width = a.raw_logits ? 1 : medfilt_width;

    long long clusters = sm_count / csize;  // one persistent CTA per SM
    if (clusters > tiles) clusters = tiles;
    if (clusters < 1) clusters = 1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(clusters * csize));
    cfg.blockDim = dim3(tc::kThreads);
    cfg.dynamicSmemBytes = tc::kSmemBytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)csize;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;

#define WCA_GO(Wv)                                                                                              \
    do {                                                                                                        \
        WCA_CUDA(cudaFuncSetAttribute(tc::capture_tc_kernel<Wv>, cudaFuncAttributeMaxDynamicSharedMemorySize,    \
                                      tc::kSmemBytes));                                                         \
        WCA_CUDA(cudaLaunchKernelEx(&cfg, tc::capture_tc_kernel<Wv>, maps, a));                                 \
    } while (0)
    switch (width) {
        case 1: WCA_GO(1); break;
        case 3: WCA_GO(3); break;
        case 5: WCA_GO(5); break;
        case 7: WCA_GO(7); break;
        default:
            set_error("capture_tc: medfilt_width=%d has no tensor-core instantiation", width);
            return WCA_ERR_UNSUPPORTED;
    }
#undef WCA_GO
    WCA_LAUNCH_CHECK("capture_tc_kernel");
    return WCA_OK;
}

}  // namespace wca
