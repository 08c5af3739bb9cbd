// tcgen05 / TMEM / TMA cross-attention capture (north-star kernel 1).
//
// Replaces reference timing.py:50-66 + upstream qkv_attention (SDPA off): for one
// (utterance, layer, head, block of 128 tokens) a CTA -- or a cluster of CTAs when the
// utterance has more frames than one CTA's tensor memory holds -- computes
//     P[t, f] = softmax_f( median_w( (q_t * s) . (k_f * s) )[:F] * qk_scale ),  s = 64^-1/4
// and writes it ONCE, already trimmed to F, in the (L, H, T, F) layout get_attentions
// returns.  The 1500-frame logits, the torch.cat copy, the reflect-padded copy and the two
// softmax passes of the reference never touch HBM.
//
// Mapping
//   tokens  <-> TMEM lanes (UMMA M = 128): thread r of an epilogue warp owns token row r, so
//               the median window and both softmax reductions are private to a thread;
//   frames  <-> TMEM columns: the fp32 accumulator IS the row buffer (filter and exp results
//               are written back in place with tcgen05.st), 464 own frames + 2 x 16 halo
//               columns per CTA;
//   F > 464 : thread-block cluster of 2/4/8 CTAs along frames; per-row max and sum are
//               exchanged through distributed shared memory (2 cluster barriers per tile);
//   fp32 in, fp32-grade out on the tensor pipe: every operand is split v = hi + lo with
//               hi = tf32-truncated v, and hi*hi + hi*lo + lo*hi is accumulated (3 x 8
//               tcgen05.mma kind::tf32 per 64-frame chunk; |error| ~ 2^-21 relative);
//   TMA     : Q / K rows arrive by cp.async.bulk (256 B per head row) on an mbarrier into a
//               padded staging tile; converter threads split them into the no-swizzle
//               K-major UMMA layout [k-chunk][8-row group][8 rows][16 B].
//
// v1 schedule: load/convert/MMA are pipelined against each other per 64-frame chunk; the
// three epilogue sweeps start when the last MMA has committed.
#include "common.cuh"

namespace wca {

namespace tc {

constexpr int kThreads = 256;
constexpr int kRows = 128;           // UMMA M: tokens per tile
constexpr int kChunk = 64;           // frames per MMA group (UMMA N of a full chunk)
constexpr int kHalo = 16;            // halo columns on either side (>= WCA max half-width 15)
constexpr int kOwnCol0 = 16;         // TMEM column of a CTA's first own frame
constexpr int kMaxOwn = 464;         // own frames per CTA: 16 + 464 + 16 <= 496 columns
constexpr int kTmemCols = 512;
constexpr int kRowBytes = kHeadDim * 4;          // 256
constexpr int kStagePitch = kRowBytes + 16;      // padded: conflict-free 16-byte column reads
constexpr int kStageRows = 64;
constexpr int kStageBytes = kStageRows * kStagePitch;   // 17408
constexpr int kQSplitBytes = kRows * kRowBytes;         // 32768 per hi / lo
constexpr int kKSplitBytes = kChunk * kRowBytes;        // 16384 per hi / lo per buffer
constexpr uint32_t kLboQ = kRows * 16;   // bytes between consecutive 16-byte k-chunks (A operand)
constexpr uint32_t kLboK = kChunk * 16;  // same for the B operand
constexpr uint32_t kSbo = 128;           // bytes between consecutive 8-row groups

constexpr int kOffStage = 0;
constexpr int kOffQHi = kOffStage + 2 * kStageBytes;
constexpr int kOffQLo = kOffQHi + kQSplitBytes;
constexpr int kOffKHi = kOffQLo + kQSplitBytes;
constexpr int kOffKLo = kOffKHi + 2 * kKSplitBytes;
constexpr int kOffStat = kOffKLo + 2 * kKSplitBytes;      // smax[128], ssum[128]
constexpr int kOffBar = kOffStat + 2 * kRows * 4;          // full[2], kfree[2], done
constexpr int kOffTmem = kOffBar + 5 * 8;
constexpr int kSmemBytes = kOffTmem + 16;
static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");
constexpr int kTilePitch = 17;  // odd pitch: conflict-free column writes
static_assert(4 * 32 * kTilePitch * 4 <= 2 * kStageBytes, "epilogue transpose tiles reuse the staging area");

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t cluster_nctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ float ld_dsmem_f32(const float *local, uint32_t rank) {
    uint32_t remote;
    float v;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(local)), "r"(rank));
    asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(remote) : "memory");
    return v;
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug must surface as a launch failure, never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    for (uint32_t spins = 0; !mbar_try_wait(bar, parity); ++spins)
        if (spins > (1u << 24)) __trap();
}
// TMA bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void tma_load_row(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]^T, both operands K-major, tf32 inputs, fp32 accumulate.
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// mbarrier arrives once every previously issued MMA of this thread has completed.
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// Shared-memory matrix descriptor, no swizzle, K-major: start address, leading-dimension
// byte offset (between 16-byte k-chunks), stride byte offset (between 8-row groups), all
// in 16-byte units; bits [46,48) = 1 selects the sm_100 descriptor version.
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((addr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
// Instruction descriptor kind::tf32: D fp32 (bits 4-5 = 1), A and B tf32 (bits 7-9, 10-12 = 2),
// both K-major (bits 15, 16 = 0), N >> 3 at bit 17, M >> 4 at bit 24.
__host__ __device__ constexpr uint32_t instr_desc_tf32(int m, int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    tmem_wait_ld();
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(
            taddr),
        "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
        "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
        "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])),
        "r"(__float_as_uint(v[11])), "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])),
        "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
        : "memory");
}
__device__ __forceinline__ float tmem_ld1(uint32_t taddr) {
    uint32_t r;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr) : "memory");
    tmem_wait_ld();
    return __uint_as_float(r);
}
__device__ __forceinline__ void tmem_st1(uint32_t taddr, float v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "r"(__float_as_uint(v)) : "memory");
}

// ------------------------------------------------------------------ building blocks
// One warp streams `n_valid` 256-byte head rows into a staging tile; bar expects the bytes.
__device__ __forceinline__ void producer_load(uint32_t stage, const float *src, int64_t ld, int n_valid, uint32_t bar,
                                              int lane) {
    if (lane == 0) mbar_expect_tx(bar, (uint32_t)n_valid * kRowBytes);
    __syncwarp();
    for (int r = lane; r < n_valid; r += kWarp)
        tma_load_row(stage + r * kStagePitch, src + (int64_t)r * ld, kRowBytes, bar);
}

// All threads: staging tile (64 rows) -> scaled hi / lo halves in UMMA no-swizzle layout.
__device__ __forceinline__ void split_stage(const unsigned char *stage, unsigned char *hi, unsigned char *lo,
                                            uint32_t lbo, int row_off, float s, int tid) {
#pragma unroll
    for (int it = 0; it < (kStageRows * 16) / kThreads; ++it) {
        const int e = it * kThreads + tid;
        const int row = e & (kStageRows - 1), ch = e >> 6;
        const float4 v = *reinterpret_cast<const float4 *>(stage + row * kStagePitch + ch * 16);
        const float x[4] = {v.x * s, v.y * s, v.z * s, v.w * s};
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
    }
    fence_proxy_async_smem();  // generic-proxy stores must be visible to the tensor core's async proxy
}

// Sliding median over a 16-column block with its neighbour blocks; all indices are compile
// time after unrolling, so everything stays in registers.
template <int W>
__device__ __forceinline__ void median_block(const float (&prev)[16], const float (&cur)[16], const float (&next)[16],
                                             float (&out)[16]) {
    constexpr int half = W / 2;
    float line[48];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        line[i] = prev[i];
        line[16 + i] = cur[i];
        line[32 + i] = next[i];
    }
#pragma unroll
    for (int pos = 0; pos < 16; ++pos) {
        if constexpr (W == 1) {
            out[pos] = cur[pos];
        } else {
            float win[W];
#pragma unroll
            for (int j = 0; j < W; ++j) win[j] = line[16 + pos - half + j];
            out[pos] = median_regs<W>(win);
        }
    }
}

// ------------------------------------------------------------------ the kernel
template <int W>
__global__ void __launch_bounds__(kThreads, 1)
capture_tc_kernel(const __grid_constant__ LayerPtrs ptrs, const wca_utt_t *__restrict__ utts, int n_heads,
                  int lh_count, int tok_blocks, int64_t ld_q, int64_t ld_k, float s, float qk_scale, int raw_logits,
                  float *__restrict__ ws) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t crank = cluster_ctarank(), csize = cluster_nctarank();

    // ---- tile decode (identical for every CTA of a cluster) ----------------------------
    const int tile = blockIdx.x / csize;
    const int tb = tile % tok_blocks;
    const int lh = (tile / tok_blocks) % lh_count;
    const int ub = tile / (tok_blocks * lh_count);
    const wca_utt_t u = utts[ub];
    const int T = u.n_tokens, F = u.n_frames;
    const int t0 = tb * kRows;
    if (t0 >= T) return;
    const int rows_valid = min(kRows, T - t0);
    const int layer = lh / n_heads, head = lh % n_heads;
    const float *qsrc = ptrs.q[layer] + (u.q_row0 + t0) * ld_q + (int64_t)head * kHeadDim;
    const float *ksrc = ptrs.k[layer] + u.k_row0 * ld_k + (int64_t)head * kHeadDim;
    float *out = ws + u.ws_off + ((int64_t)lh * T + t0) * F;

    // frame slab of this CTA
    const int slab = (((F + (int)csize - 1) / (int)csize) + 15) & ~15;
    const int f0 = (int)crank * slab;
    const int f1 = min(F, f0 + slab);
    const int n_own = max(0, f1 - f0);
    const bool filter = !raw_logits && W > 1 && F > W / 2;  // identity for very short rows, like upstream
    const int half = filter ? W / 2 : 0;
    const int m0 = f0 > 0 ? f0 - kHalo : 0;                 // first frame the MMA computes
    const int mcol0 = f0 > 0 ? 0 : kOwnCol0;                // its TMEM column: frame f sits at column f - f0 + 16
    const int m1 = n_own > 0 ? min(f1 + kHalo, F) : m0;
    const int n_mma = m1 - m0;
    const int n_chunks = (n_mma + kChunk - 1) / kChunk;

    float *smax = reinterpret_cast<float *>(smem + kOffStat);
    float *ssum = smax + kRows;
    const uint32_t bar_full0 = smem_u32(smem + kOffBar), bar_full1 = bar_full0 + 8;
    const uint32_t bar_kfree0 = bar_full0 + 16, bar_kfree1 = bar_full0 + 24, bar_done = bar_full0 + 32;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + kOffTmem);
    const uint32_t stage0 = smem_u32(smem + kOffStage), stage1 = stage0 + kStageBytes;

    // ---- setup ------------------------------------------------------------------------
    if (tid == 0) {
        mbar_init(bar_full0, 1);
        mbar_init(bar_full1, 1);
        mbar_init(bar_kfree0, 1);
        mbar_init(bar_kfree1, 1);
        mbar_init(bar_done, 1);
        fence_mbar_init();
    }
    if (tid < kRows) {
        smax[tid] = -INFINITY;
        ssum[tid] = 0.f;
    }
    if (warp == 0) {
        tmem_alloc(smem_u32(tmem_slot), kTmemCols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (n_own > 0) {
        // ---- Q: two 64-row halves through the two staging tiles ------------------------
        if (warp == 4) {
            producer_load(stage0, qsrc, ld_q, min(rows_valid, kStageRows), bar_full0, lane);
            producer_load(stage1, qsrc + (int64_t)kStageRows * ld_q, ld_q, max(rows_valid - kStageRows, 0), bar_full1, lane);
        }
        mbar_wait(bar_full0, 0);
        split_stage(smem + kOffStage, smem + kOffQHi, smem + kOffQLo, kLboQ, 0, s, tid);
        mbar_wait(bar_full1, 0);
        split_stage(smem + kOffStage + kStageBytes, smem + kOffQHi, smem + kOffQLo, kLboQ, kStageRows, s, tid);
        __syncthreads();

        // ---- K chunks: TMA(j+2) | split(j+1) | MMA(j) ----------------------------------
        auto chunk_rows = [&](int j) { return min(kChunk, n_mma - j * kChunk); };
        if (warp == 4) {
            producer_load(stage0, ksrc + (int64_t)m0 * ld_k, ld_k, chunk_rows(0), bar_full0, lane);
            if (n_chunks > 1)
                producer_load(stage1, ksrc + (int64_t)(m0 + kChunk) * ld_k, ld_k, chunk_rows(1), bar_full1, lane);
        }
        for (int j = 0; j < n_chunks; ++j) {
            const int sidx = j & 1;
            const uint32_t bar_full = sidx ? bar_full1 : bar_full0;
            const uint32_t bar_kfree = sidx ? bar_kfree1 : bar_kfree0;
            mbar_wait(bar_full, (uint32_t)(1 + (j >> 1)) & 1u);             // completion #1+(j>>1) of this stage
            if (j >= 2) mbar_wait(bar_kfree, (uint32_t)((j >> 1) - 1) & 1u);  // MMA j-2 released this K buffer
            unsigned char *khi = smem + kOffKHi + sidx * kKSplitBytes;
            unsigned char *klo = smem + kOffKLo + sidx * kKSplitBytes;
            split_stage(smem + kOffStage + sidx * kStageBytes, khi, klo, kLboK, 0, s, tid);
            tc_fence_before();
            __syncthreads();  // splits visible, staging tile `sidx` free again
            if (warp == 4 && j + 2 < n_chunks)
                producer_load(sidx ? stage1 : stage0, ksrc + (int64_t)(m0 + (j + 2) * kChunk) * ld_k, ld_k,
                              chunk_rows(j + 2), bar_full, lane);
            if (tid == 32 * 5) {  // one thread issues the MMAs of this chunk
                tc_fence_after();
                const int n_cols = (chunk_rows(j) + 15) & ~15;  // UMMA N: multiple of 16 for M = 128
                const uint32_t idesc = instr_desc_tf32(kRows, n_cols);
                const uint32_t d = tmem_base + (uint32_t)(mcol0 + j * kChunk);
                const uint32_t a_hi = smem_u32(smem + kOffQHi), a_lo = smem_u32(smem + kOffQLo);
                const uint32_t b_hi = smem_u32(khi), b_lo = smem_u32(klo);
                // small terms first: lo*hi, hi*lo, then hi*hi
#pragma unroll
                for (int pass = 0; pass < 3; ++pass) {
                    const uint32_t a = pass == 0 ? a_lo : a_hi;
                    const uint32_t b = pass == 1 ? b_lo : b_hi;
#pragma unroll
                    for (int ks = 0; ks < kHeadDim / 8; ++ks)  // K = 8 tf32 (two 16-byte k-chunks) per instruction
                        umma_tf32(d, smem_desc(a + ks * 2 * kLboQ, kLboQ, kSbo), smem_desc(b + ks * 2 * kLboK, kLboK, kSbo),
                                  idesc, (pass | ks) != 0);
                }
                umma_commit(bar_kfree);
                if (j == n_chunks - 1) umma_commit(bar_done);
            }
        }
        mbar_wait(bar_done, 0);
        tc_fence_after();
    }
    __syncthreads();

    // ---- epilogue: warps 0-3, thread <-> token row ----------------------------------------
    const int row = tid;  // valid for tid < 128
    const bool epi_warp = warp < 4;
    const bool row_ok = epi_warp && row < rows_valid;
    const bool warp_has_rows = epi_warp && (warp * 32 < rows_valid) && n_own > 0;
    const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16);  // lane quarter of this warp
    const int n_blocks = (n_own + 15) >> 4;

    if (warp_has_rows && filter) {
        // materialise the reflect padding inside TMEM: frame -i <- frame i, frame F-1+i <- frame F-1-i
        if (f0 == 0)
            for (int i = 1; i <= half; ++i) tmem_st1(trow + (uint32_t)(kOwnCol0 - i), tmem_ld1(trow + (uint32_t)(kOwnCol0 + i)));
        for (int g = max(F, f1); g < f1 + half; ++g)
            tmem_st1(trow + (uint32_t)(g - f0 + kOwnCol0), tmem_ld1(trow + (uint32_t)(2 * (F - 1) - g - f0 + kOwnCol0)));
        tmem_wait_st();
    }

    float row_max = -INFINITY;
    if (warp_has_rows && !raw_logits) {
        // sweep A: median filter (in place), * qk_scale, running max
        float prev[16], cur[16], next[16], med[16];
        tmem_ld16(trow + (uint32_t)(kOwnCol0 - 16), prev);
        tmem_ld16(trow + (uint32_t)kOwnCol0, cur);
        for (int b = 0; b < n_blocks; ++b) {
            tmem_ld16(trow + (uint32_t)(kOwnCol0 + 16 * (b + 1)), next);
            if (filter) median_block<W>(prev, cur, next, med);
            else {
#pragma unroll
                for (int i = 0; i < 16; ++i) med[i] = cur[i];
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                med[i] *= qk_scale;
                if (16 * b + i < n_own) row_max = fmaxf(row_max, med[i]);
            }
            tmem_st16(trow + (uint32_t)(kOwnCol0 + 16 * b), med);
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                prev[i] = cur[i];
                cur[i] = next[i];
            }
        }
        tmem_wait_st();
        if (row_ok) smax[row] = row_max;
    }
    float inv_sum = 1.f;
    if (!raw_logits) {
        cluster_sync_all();
        float gmax = -INFINITY;
        if (row_ok)
            for (uint32_t r = 0; r < csize; ++r) gmax = fmaxf(gmax, ld_dsmem_f32(&smax[row], r));
        // sweep B: e = exp(x - max) in place, running sum
        float row_sum = 0.f;
        if (warp_has_rows) {
            const float kLog2e = 1.4426950408889634f;
            const float shift = gmax * kLog2e;
            float v[16];
            for (int b = 0; b < n_blocks; ++b) {
                tmem_ld16(trow + (uint32_t)(kOwnCol0 + 16 * b), v);
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    v[i] = exp2f(fmaf(v[i], kLog2e, -shift));
                    if (16 * b + i < n_own) row_sum += v[i];
                }
                tmem_st16(trow + (uint32_t)(kOwnCol0 + 16 * b), v);
            }
            tmem_wait_st();
            if (row_ok) ssum[row] = row_sum;
        }
        cluster_sync_all();
        float gsum = 0.f;
        if (row_ok)
            for (uint32_t r = 0; r < csize; ++r) gsum += ld_dsmem_f32(&ssum[row], r);
        inv_sum = 1.f / gsum;
    }

    // sweep C: normalise, transpose 32x16 blocks through shared memory, coalesced row stores
    if (warp_has_rows) {
        float *tile = reinterpret_cast<float *>(smem + kOffStage) + warp * (32 * kTilePitch);
        const int rows_here = min(32, rows_valid - warp * 32);
        float v[16];
        for (int b = 0; b < n_blocks; ++b) {
            tmem_ld16(trow + (uint32_t)(kOwnCol0 + 16 * b), v);
#pragma unroll
            for (int i = 0; i < 16; ++i) tile[lane * kTilePitch + i] = v[i] * inv_sum;
            __syncwarp();
            // lane -> (row parity, column): two rows of 16 columns per store instruction
            const int c = lane & 15, rsel = lane >> 4;
            const int f = f0 + 16 * b + c;
#pragma unroll 4
            for (int rr = 0; rr < 32; rr += 2) {
                const int r = rr + rsel;
                if (r < rows_here && f < f1) st_stream(out + (int64_t)(warp * 32 + r) * F + f, tile[r * kTilePitch + c]);
            }
            __syncwarp();
        }
    }

    // ---- teardown: nobody leaves while a peer may still read our statistics ----------------
    tc_fence_before();
    cluster_sync_all();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

}  // namespace tc

bool capture_tc_supported(int max_tokens, int max_frames, int medfilt_width) {
    (void)max_tokens;
    const bool width_ok = medfilt_width == 1 || medfilt_width == 3 || medfilt_width == 5 || medfilt_width == 7;
    return width_ok && max_frames <= 8 * tc::kMaxOwn;
}

int launch_capture_tc(const float *const *h_q_layers, const float *const *h_k_layers, int n_layers, int n_heads,
                      int64_t ld_q, int64_t ld_k, const wca_utt_t *d_utts, int n_utts, int max_tokens, int max_frames,
                      int medfilt_width, float qk_scale, float *d_ws, unsigned flags, int sm_count, cudaStream_t stream) {
    (void)sm_count;
    LayerPtrs p;
    for (int l = 0; l < WCA_MAX_LAYERS; ++l) {
        p.q[l] = l < n_layers ? h_q_layers[l] : nullptr;
        p.k[l] = l < n_layers ? h_k_layers[l] : nullptr;
    }
    int csize = 1;
    while (csize < 8 && ((((max_frames + csize - 1) / csize) + 15) & ~15) > tc::kMaxOwn) csize *= 2;
    const int tok_blocks = (max_tokens + tc::kRows - 1) / tc::kRows;
    const int lh_count = n_layers * n_heads;
    const long long tiles = (long long)n_utts * lh_count * tok_blocks;
    if (tiles * csize > 0x7fffffffLL) {
        set_error("capture_tc: grid of %lld CTAs too large", tiles * csize);
        return WCA_ERR_UNSUPPORTED;
    }
    const float s = (float)0.35355339059327373;  // 64 ** -0.25 as the reference's fp32 scalar
    const int raw = (flags & WCA_CAPTURE_RAW_LOGITS) ? 1 : 0;
    const int width = raw ? 1 : medfilt_width;

    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(tiles * csize));
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

#define WCA_GO(Wv)                                                                                                  \
    do {                                                                                                            \
        WCA_CUDA(cudaFuncSetAttribute(tc::capture_tc_kernel<Wv>, cudaFuncAttributeMaxDynamicSharedMemorySize,        \
                                      tc::kSmemBytes));                                                             \
        WCA_CUDA(cudaLaunchKernelEx(&cfg, tc::capture_tc_kernel<Wv>, p, d_utts, n_heads, lh_count, tok_blocks, ld_q, \
                                    ld_k, s, qk_scale, raw, d_ws));                                                 \
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
