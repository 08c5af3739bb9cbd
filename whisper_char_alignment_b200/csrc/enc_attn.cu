// tcgen05 / TMEM / TMA attention, unmasked or causal (the Whisper audio encoder's self-attention, the decoder's
// cross-attention OUTPUT), fp32 in / fp32-grade out.
//
// Row a2 of the scope table (the teacher-forced forward, reference timing.py:57-58): the
// reference runs the encoder's 1500 x 1500 self-attention in fp32 (upstream
// MultiHeadAttention.qkv_attention; SDPA's fp32 path is a CUDA-core kernel and is a third of a
// step once the linears run on the tensor cores).  This kernel keeps fp32 accuracy on the
// tensor pipe with the same error-compensated split the capture kernel uses:
//     x = hi + lo  (hi = x chopped to tf32 by the tensor core itself, lo = x - hi exact),
//     a.b ~= lo.hi + hi.lo + hi.hi
// for BOTH contractions (Q K^T and P V), 3 x tcgen05.mma.kind::tf32 each.
//
// One CTA = 128 queries of one (batch, head); keys are swept in blocks of 64.
//   TMEM (512 columns)  Q_hi | Q_lo | S0 | S1 | S2 | P_lo0 | P_lo1 | T      (64 columns each)
//                       Three S buffers: the dependency cycle of one buffer (Q K^T -> softmax -> P V ->
//                       next Q K^T into it) is ~3 700 cycles, so with two buffers a key block could not
//                       take less than ~1 850 cycles although its MMAs need 1 536; P_lo needs only two
//                       buffers and T one (it is folded into registers before the next P V lands).
//                       Q and P are A operands read straight from tensor memory (tcgen05.mma
//                       with [a_tmem]); P_hi overwrites S in place.  T_b = P_j V_j of ONE key block
//                       (24 MMAs); the blocks are summed in fp32 registers with round-to-nearest,
//                       because the tensor core truncates when it accumulates: a 576-MMA chain into
//                       one accumulator drifts by ~1e-5 relative, a 24-MMA chain by ~1e-6.
//   shared memory       3-stage ring of {K_hi, V_hi, K_lo, V_lo} (16 KB each), K and V halves released and
//                       refilled independently (K right after Q K^T, V after P V).  TMA lands K and V
//                       with the 128-byte swizzle; that IS the UMMA canonical layout (K-major for
//                       K in Q K^T, MN-major for V in P V), so the split is a linear sweep that
//                       leaves the tile alone (it is hi) and writes lo to a twin tile.
//   softmax             one thread per query row, online with a LAZY rescale: the running reference m
//                       only moves when a block maximum exceeds it by more than 2^32 (exp arguments
//                       stay far inside the fp32 range), so a block is normally ONE streaming pass
//                       (exp against the current reference and the block maximum together) and O / l
//                       are rescaled only in that rare case (factor handed over through shared memory).
// Roles (15 warps): warp 0 P V issuer, warp 1 Q K^T issuer (two warps on two schedulers: issuing one
// 128x64x8 MMA costs about as many cycles as the tensor pipe needs to run it), warps 2-5 operand
// splitters (also split Q into TMEM), warps 6-9 softmax, warps 10-13 accumulate (own the rows of O:
// 64 registers) and epilogue, warp 14 TMA producer.  Thread <-> TMEM lane <-> query row; a warp
// reaches TMEM lanes 32 * (warp % 4) .. + 31.
#include <cuda.h>

#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace wca {
namespace ea {

using namespace tc;

constexpr int kThreads = 480;
constexpr int kQRows = 128;                    // UMMA M
constexpr int kKeys = 64;                      // keys per block: UMMA N of Q K^T, K extent of P V
constexpr int kStages = 3;
constexpr int kBoxCols = 32;                   // floats per TMA box row = 128 B = swizzle span
constexpr int kQBoxBytes = kQRows * 128;       // 16384
constexpr int kQBytes = 2 * kQBoxBytes;        // two column halves
constexpr int kKvBoxBytes = kKeys * 128;       // 8192
constexpr int kOpBytes = 2 * kKvBoxBytes;      // one operand part (K or V, hi or lo): 16384
constexpr int kRawBytes = 2 * kOpBytes;        // what TMA writes per stage: K then V
constexpr int kStageBytes = 2 * kRawBytes;     // + the lo twins
constexpr int kOffQ = 0;
constexpr int kOffStage = kOffQ + kQBytes;
constexpr int kOffFactor = kOffStage + kStages * kStageBytes;   // [4][128] rescale factors + [128] row sums
constexpr int kOffBar = kOffFactor + 5 * kQRows * 4;
constexpr int kSBufs = 3;  // S / P_hi buffers in tensor memory
enum Bar {
    kQFull = 0,
    kQReady = 1,
    // K and V of a stage have their own barrier rings: the K half is free as soon as Q K_j^T has run, a whole
    // softmax + P V earlier than the V half, so K_{j+3} is fetched and split while block j is still in flight
    kKFull = 2,
    kVFull = kKFull + kStages,
    kKSplit = kVFull + kStages,
    kVSplit = kKSplit + kStages,
    kKEmpty = kVSplit + kStages,
    kVEmpty = kKEmpty + kStages,
    kSFull = kVEmpty + kStages,   // the per-key-block barriers below form rings of kSBufs
    kPHi = kSFull + kSBufs,        // softmax -> accumulate warpgroup: P (= P_hi) stored in S[b]
    kPReady = kPHi + kSBufs,       // accumulate warpgroup -> P V issuer: P_lo stored as well
    kPvDone = kPReady + kSBufs,    // count 1 (tcgen05.commit); waited by the Q K^T issuer, the softmax AND the accumulate warpgroup
    kTFree = kPvDone + kSBufs,     // accumulate warpgroup has read T
    kLReady = kTFree + 1,          // row sums published
    kNumBars = kLReady + 1
};
constexpr int kOffTmem = kOffBar + kNumBars * 8;
constexpr int kSmemBytes = kOffTmem + 16;
static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");

constexpr uint32_t kColQHi = 0, kColQLo = 64, kColS = 128, kColPLo = 320, kColT = 448, kTmemCols = 512;
constexpr float kRescaleGap = 32.f;  // log2 units

__device__ __forceinline__ void tma_load_box3(uint32_t dst, const CUtensorMap *map, int c0, int c1, int c2, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
        "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
        : "memory");
}
// debug timeline: CTA (0,0,0) stores the low 32 bits of clock64() per (event, key block) behind the dump area
enum Ev { kEvTmaIssue = 0, kEvKvFull, kEvSplitDone, kEvQkIssue, kEvQkIssued, kEvSFull, kEvExpDone, kEvPArrive, kEvPReady,
          kEvPvIssued, kEvPvDone, kEvFoldDone, kNumEv };
__device__ __forceinline__ void stamp(float *dbg, bool on, int ev, int j) {
    if (on && j < 32) reinterpret_cast<uint32_t *>(dbg)[17000 + ev * 32 + j] = (uint32_t)clock64();
}

struct Maps {
    // 3-D (H*64, n_ctx, batch) fp32; boxes 32 x 128 x 1 (q), 32 x 64 x 1 (k, v); q and k with the
    // 128-byte swizzle, v with the 128-byte swizzle on 32-byte atoms (MN-major tf32 operand)
    CUtensorMap q, k, v;
};
struct Args {
    float *out;
    int64_t ld_out;
    int n_q, n_ctx;    // query rows and key/value rows per batch item
    int causal;        // key j is visible to query i only for j <= i (the decoder's self-attention)
    float scale_log2;  // Dh^-1/2 * log2(e)
    unsigned skip;     // debug only (env WCA_EA_SKIP): 2 P V MMAs, 4 Q K^T MMAs, 8 operand split
    float *dbg;        // debug only (wca_debug_enc_attn_buffer): CTA (0,0,0) dumps S of block 0, raw O, l, m
};

__global__ void __launch_bounds__(kThreads, 1) enc_attn_kernel(const __grid_constant__ Maps maps, const __grid_constant__ Args a) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q0 = blockIdx.x * kQRows, head = blockIdx.y, batch = blockIdx.z;
    const int col0 = head * kHeadDim;
    // causal: the key blocks past the tile's last query row are never visited
    const int n_blocks = ((a.causal ? min(a.n_ctx, q0 + kQRows) : a.n_ctx) + kKeys - 1) / kKeys;
    const uint32_t bars = smem_u32(smem + kOffBar);
    auto bar = [&](int which) { return bars + 8u * (uint32_t)which; };
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + kOffTmem);
    // One arrival per warp: every lane has finished (and fenced) its part, the warp converges, lane 0 arrives
    // (128 threads arriving on one mbarrier word serialise on it).
    auto warp_arrive = [&](uint32_t b) {
        __syncwarp();
        if (lane == 0) mbar_arrive(b);
    };
    const bool tr = a.dbg != nullptr && blockIdx.x + blockIdx.y + blockIdx.z == 0 && lane == 0 && (warp < 2 || (warp & 3) == 2);  // incl. warp 14

    if (tid == 0) {
        mbar_init(bar(kQFull), 1);
        mbar_init(bar(kQReady), 4);
        for (int i = 0; i < kStages; ++i) {
            mbar_init(bar(kKFull + i), 1);
            mbar_init(bar(kVFull + i), 1);
            mbar_init(bar(kKSplit + i), 4);
            mbar_init(bar(kVSplit + i), 4);
            mbar_init(bar(kKEmpty + i), 1);
            mbar_init(bar(kVEmpty + i), 1);
        }
        for (int i = 0; i < kSBufs; ++i) {
            mbar_init(bar(kSFull + i), 1);
            mbar_init(bar(kPHi + i), 4);
            mbar_init(bar(kPReady + i), 4);
            mbar_init(bar(kPvDone + i), 1);
        }
        mbar_init(bar(kTFree), 4);
        mbar_init(bar(kLReady), 4);
        fence_mbar_init();
    }
    if (warp == 0) {
        tmem_alloc(smem_u32(tmem_slot), kTmemCols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 14) {
        // ================= TMA producer =================
        if (lane == 0) {
            tma_prefetch_map(&maps.q);
            tma_prefetch_map(&maps.k);
            tma_prefetch_map(&maps.v);
            mbar_expect_tx(bar(kQFull), kQBytes);
            tma_load_box3(smem_u32(smem + kOffQ), &maps.q, col0, q0, batch, bar(kQFull));
            tma_load_box3(smem_u32(smem + kOffQ) + kQBoxBytes, &maps.q, col0 + kBoxCols, q0, batch, bar(kQFull));
            // issue order K_0, K_1, V_0, K_2, V_1, ...: K runs one block ahead of V
            for (int t = 0; t <= n_blocks; ++t) {
                if (t < n_blocks) {
                    const int s = t % kStages;
                    mbar_wait(bar(kKEmpty + s), ((t / kStages) & 1) ^ 1);  // first lap passes immediately
                    stamp(a.dbg, tr, kEvTmaIssue, t);
                    const uint32_t dst = smem_u32(smem + kOffStage + s * kStageBytes);
                    mbar_expect_tx(bar(kKFull + s), kOpBytes);  // rows past n_ctx are zero-filled, boxes are always full
                    tma_load_box3(dst, &maps.k, col0, t * kKeys, batch, bar(kKFull + s));
                    tma_load_box3(dst + kKvBoxBytes, &maps.k, col0 + kBoxCols, t * kKeys, batch, bar(kKFull + s));
                }
                if (t >= 1) {
                    const int i = t - 1, s = i % kStages;
                    mbar_wait(bar(kVEmpty + s), ((i / kStages) & 1) ^ 1);
                    const uint32_t dst = smem_u32(smem + kOffStage + s * kStageBytes) + kOpBytes;
                    mbar_expect_tx(bar(kVFull + s), kOpBytes);
                    tma_load_box3(dst, &maps.v, col0, i * kKeys, batch, bar(kVFull + s));
                    tma_load_box3(dst + kKvBoxBytes, &maps.v, col0 + kBoxCols, i * kKeys, batch, bar(kVFull + s));
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ================= Q K^T issuer: the whole warp runs the loop, one elected lane issues =================
        const uint32_t idesc_qk = instr_desc_tf32(kQRows, kKeys);  // A tmem (K-major), B K-major
        const uint32_t elected = elect_one() ? 1u : 0u;
        const uint32_t tm = __shfl_sync(0xffffffffu, tmem, 0);  // warp-uniform for the compiler
        const uint32_t stage0 = smem_u32(smem + kOffStage);
        const uint32_t skip_qk = a.skip & 4;
        mbar_wait(bar(kQReady), 0);
        for (int j = 0; j < n_blocks; ++j) {
            // ---- S[j & 1] = Q K_j^T
            const int s = j % kStages, b = j % kSBufs;
            mbar_wait(bar(kKSplit + s), (j / kStages) & 1);
            if (j >= kSBufs) mbar_wait(bar(kPvDone + b), ((j - kSBufs) / kSBufs) & 1);  // P V of block j-3 no longer reads P_hi = S[b]
            tc_fence_after();
            stamp(a.dbg, tr, kEvQkIssue, j);
            const uint32_t k_hi = stage0 + (uint32_t)s * kStageBytes;
            const uint32_t d = tm + kColS + (uint32_t)b * kKeys;
            // K-major, 128B swizzle, SBO = 1024 B between 8-key groups; 8 tf32 = 32 bytes along the row,
            // the second column half is the next TMA box
            const uint64_t dk_hi = smem_desc_sw128(k_hi, 16, 1024), dk_lo = smem_desc_sw128(k_hi + kRawBytes, 16, 1024);
            if (!skip_qk) {
#pragma unroll
                for (int pass = 0; pass < 3; ++pass) {  // small terms first: lo.hi, hi.lo, hi.hi
                    const uint32_t a_col = tm + (pass == 0 ? kColQLo : kColQHi);
                    const uint64_t db = pass == 1 ? dk_lo : dk_hi;
#pragma unroll
                    for (int ks = 0; ks < kHeadDim / 8; ++ks)
                        umma_tf32_ts(d, a_col + (uint32_t)(ks * 8),
                                     db + (uint64_t)(((ks >> 2) * kKvBoxBytes + (ks & 3) * 32) >> 4), idesc_qk,
                                     (pass | ks) != 0, elected);
                }
            }
            umma_commit_if(bar(kKEmpty + s), elected);
            umma_commit_if(bar(kSFull + b), elected);
            stamp(a.dbg, tr, kEvQkIssued, j);
        }
    } else if (warp == 0) {
        // ================= P V issuer (its own warp, on another scheduler than the Q K^T issuer) =================
        const uint32_t idesc_pv = instr_desc_tf32(kQRows, kHeadDim) | (1u << 16);  // B = V, MN-major
        const uint32_t elected = elect_one() ? 1u : 0u;
        const uint32_t tm = __shfl_sync(0xffffffffu, tmem, 0);
        const uint32_t stage0 = smem_u32(smem + kOffStage);
        const uint32_t skip_pv = a.skip & 2;
        for (int i = 0; i < n_blocks; ++i) {
            // ---- T[i & 1] = P_i V_i
            const int s = i % kStages, b = i % kSBufs;
            mbar_wait(bar(kVSplit + s), (i / kStages) & 1);
            mbar_wait(bar(kPReady + b), (i / kSBufs) & 1);
            if (i >= 1) mbar_wait(bar(kTFree), (i - 1) & 1);  // T of block i-1 has been folded into the O registers
            tc_fence_after();
            stamp(a.dbg, tr, kEvPReady, i);
            const uint32_t v_hi = stage0 + (uint32_t)s * kStageBytes + kOpBytes;
            const uint32_t p_hi = tm + kColS + (uint32_t)b * kKeys, p_lo = tm + kColPLo + (uint32_t)(i & 1) * kKeys;
            // MN-major, 128B swizzle with 32B atoms: 4 keys x 128 B per atom, so the 8 keys of one MMA are
            // two atoms SBO = 512 B apart; the two 32-float halves of d are LBO apart
            const uint64_t dv_hi = smem_desc_sw128(v_hi, kKvBoxBytes, 512, kSw128Base32);
            const uint64_t dv_lo = smem_desc_sw128(v_hi + kRawBytes, kKvBoxBytes, 512, kSw128Base32);
            if (!skip_pv) {
#pragma unroll
                for (int pass = 0; pass < 3; ++pass) {
                    const uint32_t a_col = pass == 0 ? p_lo : p_hi;
                    const uint64_t db = pass == 1 ? dv_lo : dv_hi;
#pragma unroll
                    for (int ks = 0; ks < kKeys / 8; ++ks)
                        umma_tf32_ts(tm + kColT, a_col + (uint32_t)(ks * 8),
                                     db + (uint64_t)((ks * 1024) >> 4), idesc_pv, (pass | ks) != 0, elected);
                }
            }
            umma_commit_if(bar(kVEmpty + s), elected);
            umma_commit_if(bar(kPvDone + b), elected);
            stamp(a.dbg, tr, kEvPvIssued, i);
        }
    } else if (warp >= 2 && warp < 6) {
        // ================= operand splitters =================
        const int t = tid - 64;                    // linear index for the in-place K / V sweep
        const int qrow = (warp & 3) * 32 + lane;   // query row this thread splits (its TMEM lane)
        {
            mbar_wait(bar(kQFull), 0);
            const unsigned char *qs = smem + kOffQ;
            const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16);
#pragma unroll
            for (int c16 = 0; c16 < 4; ++c16) {  // 16 columns at a time
                float hi[16], lo[16];
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) {
                    const int ch = c16 * 4 + q4;  // 16-byte chunk of the 256-byte row
                    const float4 v = *reinterpret_cast<const float4 *>(qs + (ch >> 3) * kQBoxBytes + qrow * 128 +
                                                                       (((ch & 7) ^ (qrow & 7)) << 4));
                    const float x[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        hi[q4 * 4 + e] = x[e];
                        lo[q4 * 4 + e] = tf32_lo(x[e]);
                    }
                }
                tmem_st16(trow + kColQHi + (uint32_t)(c16 * 16), hi);
                tmem_st16(trow + kColQLo + (uint32_t)(c16 * 16), lo);
            }
            tmem_wait_st();
            tc_fence_before();
            warp_arrive(bar(kQReady));
        }
        // lo twin of one operand (K or V) of a stage: linear sweep, hi = the raw tile
        auto split_op = [&](int stage, int op) {
            const float4 *raw = reinterpret_cast<const float4 *>(smem + kOffStage + stage * kStageBytes + op * kOpBytes);
            float4 *twin = reinterpret_cast<float4 *>(smem + kOffStage + stage * kStageBytes + kRawBytes + op * kOpBytes);
            constexpr int kIters = kOpBytes / 16 / 128;  // 8 float4 per thread
            float4 v[kIters];
#pragma unroll
            for (int it = 0; it < kIters; ++it) v[it] = raw[it * 128 + t];
#pragma unroll
            for (int it = 0; it < kIters; ++it)
                twin[it * 128 + t] = make_float4(tf32_lo(v[it].x), tf32_lo(v[it].y), tf32_lo(v[it].z), tf32_lo(v[it].w));
            fence_proxy_async_smem();  // generic-proxy stores -> visible to the tensor core's async proxy
        };
        for (int tt = 0; tt <= n_blocks; ++tt) {  // same order as the producer: K_0, K_1, V_0, K_2, V_1, ...
            if (tt < n_blocks) {
                const int s = tt % kStages;
                mbar_wait(bar(kKFull + s), (tt / kStages) & 1);
                stamp(a.dbg, tr, kEvKvFull, tt);
                if (!(a.skip & 8)) split_op(s, 0);
                warp_arrive(bar(kKSplit + s));
                stamp(a.dbg, tr, kEvSplitDone, tt);
            }
            if (tt >= 1) {
                const int i = tt - 1, s = i % kStages;
                mbar_wait(bar(kVFull + s), (i / kStages) & 1);
                if (!(a.skip & 8)) split_op(s, 1);
                warp_arrive(bar(kVSplit + s));
            }
        }
    } else if (warp >= 6 && warp < 10) {
        // ================= softmax: thread <-> query row =================
        const int row = (warp & 3) * 32 + lane;
        const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16);
        const float c = a.scale_log2;
        float *factor = reinterpret_cast<float *>(smem + kOffFactor);  // [4][128] rescale factors, ring over blocks
        float m_ref = 0.f, l_sum = 0.f;
        for (int j = 0; j < n_blocks; ++j) {
            const int b = j % kSBufs;
            const int n_block = min(kKeys, a.n_ctx - j * kKeys);  // < 64 only for the last block
            // keys of this block the thread's query row may see (causal: up to its own position; <= 0: none)
            const int n_valid = a.causal ? min(n_block, q0 + row + 1 - j * kKeys) : n_block;
            mbar_wait(bar(kSFull + b), (j / kSBufs) & 1);
            tc_fence_after();
            stamp(a.dbg, tr, kEvSFull, j);
            const uint32_t s_col = trow + kColS + (uint32_t)b * kKeys;
            float sv[4][16];
#pragma unroll
            for (int q = 0; q < 4; ++q) tmem_ld16_issue(s_col + (uint32_t)(q * 16), sv[q]);
            if (a.dbg && j == 0 && blockIdx.x + blockIdx.y + blockIdx.z == 0) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    tmem_ld_wait(sv[q]);
#pragma unroll
                    for (int e = 0; e < 16; ++e) a.dbg[row * 64 + q * 16 + e] = sv[q][e];
                }
            }
            // Streaming pass against the CURRENT reference: p = 2^(s c - m_ref) and the block maximum are
            // computed together, chunk by chunk as the tensor-memory loads land (four independent chains).
            float bmax[4], lpart[4];
            if (j == 0) {
                // no reference yet: take the maximum of the first 16 keys.  Anything within 2^32 of the row maximum
                // will do (the lazy rescale below catches the rest), and it keeps block 0 on the hot path.
                tmem_ld_wait(sv[0]);
                float mx = sv[0][0];
#pragma unroll
                for (int e = 1; e < 16; ++e) mx = (n_valid == kKeys || e < n_valid) ? fmaxf(mx, sv[0][e]) : mx;
                m_ref = mx * c;
            }
            const float neg_m = -m_ref;
            // warp-uniform: every row of the tile sees all 64 keys (ragged: the last block of a context that is not a
            // multiple of 64; causal: the blocks the diagonal of this tile runs through)
            const bool full = n_block == kKeys && (!a.causal || j * kKeys + kKeys - 1 <= q0);
            if (full) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    tmem_ld_wait(sv[q]);
                    float mx = sv[q][0], ls = 0.f;
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        mx = fmaxf(mx, sv[q][e]);
                        sv[q][e] = ex2_approx(fmaf(sv[q][e], c, neg_m));
                        ls += sv[q][e];
                    }
                    bmax[q] = mx;
                    lpart[q] = ls;
                }
            } else {  // ragged block: only its maximum here, the masked exponentials come from the generic path below
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    tmem_ld_wait(sv[q]);
                    float mx = -INFINITY;
#pragma unroll
                    for (int e = 0; e < 16; ++e) mx = (q * 16 + e < n_valid) ? fmaxf(mx, sv[q][e]) : mx;
                    bmax[q] = mx;
                    lpart[q] = 0.f;
                }
            }
            const float bm = fmaxf(fmaxf(bmax[0], bmax[1]), fmaxf(bmax[2], bmax[3])) * c;
            float f = 1.f;
            // The reference moves only when a block exceeds it by 2^32 (rare); the same generic path serves ragged blocks.
            const bool need = bm > m_ref + kRescaleGap;
            if (__any_sync(0xffffffffu, need) || !full) {
                if (need) {
                    f = j == 0 ? 1.f : ex2_approx(m_ref - bm);  // nothing accumulated yet in block 0
                    m_ref = bm;
                    l_sum *= f;
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) tmem_ld16_issue(s_col + (uint32_t)(q * 16), sv[q]);  // S is still intact
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    tmem_ld_wait(sv[q]);
                    float ls = 0.f;
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        const bool ok = q * 16 + e < n_valid;
                        sv[q][e] = ok ? ex2_approx(fmaf(sv[q][e], c, -m_ref)) : 0.f;
                        ls += sv[q][e];
                    }
                    lpart[q] = ls;
                }
            }
            l_sum += (lpart[0] + lpart[1]) + (lpart[2] + lpart[3]);
            stamp(a.dbg, tr, kEvExpDone, j);
            factor[(j & 3) * kQRows + row] = f;  // read by the accumulate thread of this row after P V of block j
            // P is stored as it is (the tensor core chops it to P_hi); the accumulate warpgroup derives P_lo from it,
            // so this warpgroup -- the one every key block has to pass through -- is done a store round trip earlier
#pragma unroll
            for (int q = 0; q < 4; ++q) tmem_st16(s_col + (uint32_t)(q * 16), sv[q]);
            tmem_wait_st();
            tc_fence_before();
            warp_arrive(bar(kPHi + b));
            stamp(a.dbg, tr, kEvPArrive, j);
        }
        factor[4 * kQRows + row] = l_sum;
        warp_arrive(bar(kLReady));
    } else if (warp >= 10) {
        // ================= accumulate + epilogue: thread <-> query row =================
        // acc = this row of O, summed over key blocks in fp32 registers with round-to-nearest.
        const int row = (warp & 3) * 32 + lane;
        const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16);
        const float *factor = reinterpret_cast<const float *>(smem + kOffFactor);
        float acc[4][16];
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int e = 0; e < 16; ++e) acc[q][e] = 0.f;
        // acc += T once P_i V_i has landed (T is single: P V of block i+1 waits for this)
        auto fold = [&](int i) {
            mbar_wait(bar(kPvDone + i % kSBufs), (i / kSBufs) & 1);
            tc_fence_after();
            stamp(a.dbg, tr, kEvPvDone, i);
            float t[2][16];  // two chunks in flight
            const uint32_t t_col = trow + kColT;
            tmem_ld16_issue(t_col, t[0]);
            const float f = *reinterpret_cast<const volatile float *>(&factor[(i & 3) * kQRows + row]);
            if (f != 1.f) {  // the reference moved at block i: everything before it is rescaled
#pragma unroll
                for (int q = 0; q < 4; ++q)
#pragma unroll
                    for (int e = 0; e < 16; ++e) acc[q][e] *= f;
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (q + 1 < 4) tmem_ld16_issue(t_col + (uint32_t)((q + 1) * 16), t[(q + 1) & 1]);
                tmem_ld_wait(t[q & 1]);
#pragma unroll
                for (int e = 0; e < 16; ++e) acc[q][e] += t[q & 1][e];
            }
            tc_fence_before();
            warp_arrive(bar(kTFree));
            stamp(a.dbg, tr, kEvFoldDone, i);
        };
        for (int j = 0; j < n_blocks; ++j) {
            // P_lo of block j from the P the softmax warpgroup left in S[b]; P_lo[j & 1] is free: the fold of block
            // j-2 (previous iteration) waited for its P V.  Done BEFORE folding block j-1 so that it overlaps P V_{j-1}.
            const int b = j % kSBufs;
            mbar_wait(bar(kPHi + b), (j / kSBufs) & 1);
            tc_fence_after();
            const uint32_t p_col = trow + kColS + (uint32_t)b * kKeys;
            const uint32_t lo_col = trow + kColPLo + (uint32_t)(j & 1) * kKeys;
            float pv[2][16];
            tmem_ld16_issue(p_col, pv[0]);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (q + 1 < 4) tmem_ld16_issue(p_col + (uint32_t)((q + 1) * 16), pv[(q + 1) & 1]);
                tmem_ld_wait(pv[q & 1]);
                float lo[16];
#pragma unroll
                for (int e = 0; e < 16; ++e) lo[e] = tf32_lo(pv[q & 1][e]);
                tmem_st16(lo_col + (uint32_t)(q * 16), lo);
            }
            tmem_wait_st();
            tc_fence_before();
            warp_arrive(bar(kPReady + b));
            if (j >= 1) fold(j - 1);
        }
        fold(n_blocks - 1);
        // ---- epilogue: O / l -> global
        mbar_wait(bar(kLReady), 0);
        const float l_sum = *reinterpret_cast<const volatile float *>(&factor[4 * kQRows + row]);
        const float inv = 1.f / l_sum;
        if (a.dbg && blockIdx.x + blockIdx.y + blockIdx.z == 0) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int e = 0; e < 16; ++e) a.dbg[8192 + row * 64 + q * 16 + e] = acc[q][e];
            a.dbg[16384 + row] = l_sum;
        }
        // A thread owns a row, but 32 lanes storing 16 bytes of 32 different rows is 32 memory requests per
        // instruction.  The rows go through the Q staging tile (free since Q was split into tensor memory): each warp
        // writes its 32 rows with the 16-byte chunks XOR-swizzled by the row, then stores two full 256-byte rows per
        // instruction.
        float4 *stage = reinterpret_cast<float4 *>(smem + kOffQ) + (warp & 3) * 32 * 16;  // [32 rows][16 chunks]
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int e = 0; e < 16; e += 4)
                stage[lane * 16 + ((q * 4 + e / 4) ^ (lane & 15))] =
                    make_float4(acc[q][e] * inv, acc[q][e + 1] * inv, acc[q][e + 2] * inv, acc[q][e + 3] * inv);
        __syncwarp();
        {
            const int ch = lane & 15, sub = lane >> 4;
#pragma unroll 4
            for (int i = 0; i < 16; ++i) {
                const int r = 2 * i + sub;                 // row within this warp's 32
                const int grow = q0 + (warp & 3) * 32 + r;  // query row
                const float4 v = stage[r * 16 + (ch ^ (r & 15))];
                if (grow < a.n_q)
                    *reinterpret_cast<float4 *>(a.out + ((int64_t)batch * a.n_q + grow) * a.ld_out + col0 + ch * 4) = v;
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem, kTmemCols);
    }
}

}  // namespace ea

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int encode_map3(EncodeTiledFn encode, CUtensorMap *map, const float *base, int64_t cols, int64_t n_ctx, int64_t batch,
                       int64_t ld, int box_rows, CUtensorMapSwizzle swizzle) {
    const cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)n_ctx, (cuuint64_t)batch};
    const cuuint64_t strides[2] = {(cuuint64_t)ld * sizeof(float), (cuuint64_t)n_ctx * (cuuint64_t)ld * sizeof(float)};
    const cuuint32_t box[3] = {(cuuint32_t)ea::kBoxCols, (cuuint32_t)box_rows, 1};
    const cuuint32_t elem[3] = {1, 1, 1};
    const CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float *>(base), dims, strides, box, elem,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled (3-D) failed with CUresult %d (cols %lld, n_ctx %lld, batch %lld, ld %lld)", (int)r,
                  (long long)cols, (long long)n_ctx, (long long)batch, (long long)ld);
        return WCA_ERR_CUDA;
    }
    return WCA_OK;
}

static float *g_enc_attn_dbg = nullptr;
void set_enc_attn_debug_buffer(float *d_buf) { g_enc_attn_dbg = d_buf; }

int launch_full_attention(const float *d_q, const float *d_k, const float *d_v, float *d_out, int n_batch, int n_q, int n_ctx,
                          int n_heads, int64_t ld_q, int64_t ld_k, int64_t ld_v, int64_t ld_out, int causal, cudaStream_t stream) {
    static EncodeTiledFn encode = nullptr;
    if (!encode) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        WCA_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        if (qres != cudaDriverEntryPointSuccess || !fn) {
            set_error("full_attention: the driver does not export cuTensorMapEncodeTiled");
            return WCA_ERR_CUDA;
        }
        encode = reinterpret_cast<EncodeTiledFn>(fn);
    }
    ea::Maps maps;
    memset(&maps, 0, sizeof(maps));
    const int64_t cols = (int64_t)n_heads * kHeadDim;
    int rc = encode_map3(encode, &maps.q, d_q, cols, n_q, n_batch, ld_q, ea::kQRows, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    rc = encode_map3(encode, &maps.k, d_k, cols, n_ctx, n_batch, ld_k, ea::kKeys, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    rc = encode_map3(encode, &maps.v, d_v, cols, n_ctx, n_batch, ld_v, ea::kKeys, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
    if (rc) return rc;
    ea::Args a;
    a.out = d_out;
    a.ld_out = ld_out;
    a.n_q = n_q;
    a.n_ctx = n_ctx;
    a.causal = causal;
    a.dbg = g_enc_attn_dbg;
    {
        static int skip = -1;
        if (skip < 0) {
            const char *e = getenv("WCA_EA_SKIP");
            skip = e ? atoi(e) : 0;
        }
        a.skip = (unsigned)skip;
    }
    a.scale_log2 = (float)(0.125 * 1.4426950408889634);  // Dh^-1/2 (= (Dh^-1/4)^2 of upstream) * log2(e)
    WCA_CUDA(cudaFuncSetAttribute(ea::enc_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ea::kSmemBytes));
    const dim3 grid((unsigned)((n_q + ea::kQRows - 1) / ea::kQRows), (unsigned)n_heads, (unsigned)n_batch);
    ea::enc_attn_kernel<<<grid, ea::kThreads, ea::kSmemBytes, stream>>>(maps, a);
    WCA_LAUNCH_CHECK("enc_attn_kernel");
    return WCA_OK;
}

}  // namespace wca
