// tc_ptx.cuh -- thin PTX wrappers shared by the tcgen05 kernels (capture_tc.cu, enc_attn.cu):
// mbarriers, TMA tensor-map loads, TMEM allocation / load / store, tcgen05.mma (tf32) and
// its shared-memory / instruction descriptors.  sm_100a only.
#pragma once
#include <cuda.h>  // CUtensorMap

#include <cstdint>

namespace wca {
namespace tc {

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
__device__ __forceinline__ uint32_t map_to_rank(uint32_t local_smem, uint32_t rank) {
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_smem), "r"(rank));
    return remote;
}
__device__ __forceinline__ float ld_dsmem_f32(const float *local, uint32_t rank) {
    float v;
    asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(map_to_rank(smem_u32(local), rank)) : "memory");
    return v;
}

// {m, sum} pair of one row from a peer CTA.  Volatile (stays behind the caller's mbarrier wait) but without a memory
// clobber and without data dependences between ranks: the loads of all ranks are in flight together, one
// distributed-shared-memory round trip instead of one per rank.
__device__ __forceinline__ float2 ld_dsmem_f32x2(const float2 *local, uint32_t rank) {
    float2 v;
    asm volatile("ld.shared::cluster.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(map_to_rank(smem_u32(local), rank)));
    return v;
}

__device__ __forceinline__ float4 ld_dsmem_f32x4(const float4 *local, uint32_t rank) {
    float4 v;
    asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "r"(map_to_rank(smem_u32(local), rank)));
    return v;
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t local_bar, uint32_t rank) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(map_to_rank(local_bar, rank))
                 : "memory");
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
__device__ __forceinline__ bool mbar_try_wait_cluster(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded waits: a protocol bug must surface as a launch failure, never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    for (uint32_t spins = 0; !mbar_try_wait(bar, parity); ++spins)
        if (spins > (1u << 24)) __trap();
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
    for (uint32_t spins = 0; !mbar_try_wait_cluster(bar, parity); ++spins)
        if (spins > (1u << 24)) __trap();
}
// TMA tensor-map box load global -> shared, completion signalled on an mbarrier (SASS: UTMALDG).
__device__ __forceinline__ void tma_load_box(uint32_t dst, const CUtensorMap *map, int col, int row, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::
                     "r"(dst), "l"(map), "r"(col), "r"(row), "r"(bar)
                 : "memory");
}
// Same box, but only pulled into L2 (no shared-memory destination, no completion signal).
__device__ __forceinline__ void tma_prefetch_box(const CUtensorMap *map, int col, int row) {
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(map), "r"(col), "r"(row) : "memory");
}
__device__ __forceinline__ void tma_prefetch_map(const CUtensorMap *map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

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

// Asynchronous TMEM load of 16 consecutive columns of this thread's lane; the registers
// are only valid after tmem_ld_wait(), which names them so no use can be hoisted above it.
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, float (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]), "=f"(v[8]),
          "=f"(v[9]), "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait(float (&v)[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+f"(v[0]), "+f"(v[1]), "+f"(v[2]), "+f"(v[3]), "+f"(v[4]), "+f"(v[5]), "+f"(v[6]), "+f"(v[7]),
                   "+f"(v[8]), "+f"(v[9]), "+f"(v[10]), "+f"(v[11]), "+f"(v[12]), "+f"(v[13]), "+f"(v[14]), "+f"(v[15])
                 :
                 : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(
            taddr),
        "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]), "f"(v[8]), "f"(v[9]),
        "f"(v[10]), "f"(v[11]), "f"(v[12]), "f"(v[13]), "f"(v[14]), "f"(v[15])
        : "memory");
}
// four-column variants (edges of the median window, rare in-place rescale)
__device__ __forceinline__ void tmem_ld4_issue(uint32_t taddr, float (&v)[4]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_ld_wait4(float (&v)[4]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;" : "+f"(v[0]), "+f"(v[1]), "+f"(v[2]), "+f"(v[3]) : : "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const float (&v)[4]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "f"(v[0]), "f"(v[1]), "f"(v[2]),
                 "f"(v[3])
                 : "memory");
}
__device__ __forceinline__ float tmem_ld1(uint32_t taddr) {
    float r;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];\n\ttcgen05.wait::ld.sync.aligned;"
                 : "=f"(r)
                 : "r"(taddr)
                 : "memory");
    return r;
}
__device__ __forceinline__ void tmem_st1(uint32_t taddr, float v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "f"(v) : "memory");
}
// Exactly one lane of a converged warp gets true; keeps the surrounding code warp-uniform so
// descriptors live in uniform registers instead of being broadcast lane by lane.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// 128-byte-swizzled operand descriptor, sm_100 version bit 46.  Layout type (bits 61-63):
// 2 = SWIZZLE_128B (16-byte chunks XOR row & 7; what K-major tf32 operands use),
// 1 = SWIZZLE_128B_BASE32B (32-byte chunks XOR row & 3; the ONLY swizzle MN-major tf32 operands
//     accept -- TMA produces it with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B).
constexpr uint64_t kSw128 = 2, kSw128Base32 = 1;
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t addr, uint32_t lbo, uint32_t sbo, uint64_t layout = kSw128) {
    return (uint64_t)((addr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46) | (layout << 61);
}
// D[tmem] (+)= A[tmem] * B[smem]: A is read from tensor memory (lane = row, one tf32 per column).
// Executed by the WHOLE warp with `elected` true in one lane: only the instruction is predicated, so
// the descriptor arithmetic around it stays in convergent code and on the uniform datapath (with the
// loop inside a divergent `if (elected)` every operand went through R2UR: ~45 cycles per MMA issued,
// more than the 32 cycles a 128x64x8 tf32 MMA occupies the tensor pipe).
__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate, uint32_t elected) {
    asm volatile(
        "{\n\t"
        ".reg .pred p, e;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "setp.ne.b32 e, %5, 0;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
        "}" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(elected)
        : "memory");
}
__device__ __forceinline__ void umma_commit_if(uint32_t bar, uint32_t elected) {
    asm volatile(
        "{\n\t"
        ".reg .pred e;\n\t"
        "setp.ne.b32 e, %1, 0;\n\t"
        "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t"
        "}" ::"r"(bar),
        "r"(elected)
        : "memory");
}
// x = hi + lo for the 3 x tf32 product.  tcgen05.mma.kind::tf32 ignores the 13 low mantissa bits of
// its operands (measured: with hi = x left as is and lo = x - chop(x) the results are fp32-grade; a
// rounding tensor core would be off by a tf32 ulp), so hi needs no instruction and no store at all;
// lo is exact in fp32 and is chopped to its 11 leading bits by the tensor core: |error| <= 2^-21 |x|.
__device__ __forceinline__ float tf32_lo(float x) { return x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

// Same with both operands in shared memory.
__device__ __forceinline__ void umma_tf32_ss_if(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                                uint32_t accumulate, uint32_t elected) {
    asm volatile(
        "{\n\t"
        ".reg .pred p, e;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "setp.ne.b32 e, %5, 0;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(elected)
        : "memory");
}

}  // namespace tc
}  // namespace wca
