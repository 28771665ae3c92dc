// Minimal tcgen05 / TMEM / mbarrier helpers for the tensor-core shared-MLP kernels (sm_100a).
//
// Operand convention used by every kernel here: both operands are staged by the CTA's own threads
// (they need a per-element transform, so TMA cannot deliver them) into shared memory in the
// canonical K-MAJOR, NO-SWIZZLE UMMA layout for 32-bit elements: 8 x 16-byte core matrices,
//     offset(row, k) = (k / 4) * LBO + (row / 8) * SBO + (row % 8) * 16 + (k % 4) * 4
// with SBO = 128 (the 8-row groups of one k-column block are contiguous) and LBO = rows * 16.
// One tcgen05.mma.kind::tf32 consumes K = 8, i.e. two k-column blocks; the descriptor's start
// address advances by 2 * LBO per k-step.
//
// f32 accuracy on the tf32 pipe: every operand element v is split as hi = tf32(v), lo = tf32(v - hi)
// (both round-to-nearest) and each product is evaluated as hi*hi + lo*hi + hi*lo (three MMAs into
// the same f32 TMEM accumulator); the dropped lo*lo term is below 2^-22 relative and unbiased.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace i2p {
namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// cute::UMMA::SmemDescriptor (cute/arch/mma_sm100_desc.hpp): start >> 4 in [0,14), LBO >> 4 in [16,30),
// SBO >> 4 in [32,46), version = 1 in [46,48), layout_type (swizzle) = 0 in [61,64).
// layout_type: 0 no swizzle, 1 SWIZZLE_128B_BASE32B (the only layout an MN-major 32-bit operand may use), 2 128B, 4 64B, 6 32B
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type = 0) {
    uint64_t d = (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(layout_type & 7u) << 61;
    return d;
}

// cute::UMMA::InstrDescriptor for kind::tf32, f32 accumulate, both operands K-major, M = 128.
__host__ __device__ constexpr uint32_t instr_desc_tf32(int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
        :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"((uint32_t)accumulate) : "memory");
}

__device__ __forceinline__ void commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t ncols) {  // one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // the same warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(taddr), "r"(ncols) : "memory");
}

__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// generic-proxy shared-memory writes -> visible to the async proxy (the tensor core reads smem through it)
__device__ __forceinline__ void fence_smem_to_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}\n"
            : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    } while (!done);
}

// transaction-count arm + 1-D bulk copy global -> shared (the TMA engine without a tensor map): the bytes land
// through the async proxy and complete on the mbarrier, so a tcgen05.mma that is issued after waiting on it
// needs no proxy fence
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst_smem), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// 16 consecutive f32 columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ uint32_t to_tf32(float v) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return r;
}
// v = hi + lo with hi = v ROUNDED to tf32 (10 explicit mantissa bits) and lo = v - hi exactly (f32,
// signed, |lo| <= 2^-11 |v|).  The tensor core reads only the top 19 bits of a tf32 operand, i.e. it
// truncates; adding half a tf32 ulp to the bit pattern first turns that truncation into
// round-to-nearest for both parts.  Rounding (not truncating) matters: truncation errors all have
// the sign of v, so over a K-long dot product they add up linearly (measured 2.5e-6 relative at
// K = 262) instead of as a random walk; with rounding the split error (<= 2^-22 |v|) and the
// dropped lo*lo term are unbiased and the result is as close to f64 as an f32 FMA chain.
// Integer adds on the bit pattern: cvt.rna.tf32.f32 lowers to a ~10-instruction sequence on sm_100a.
__device__ __forceinline__ void split_tf32(float v, uint32_t &hi, uint32_t &lo) {
    hi = (__float_as_uint(v) + 0x1000u) & 0xffffe000u;
    lo = __float_as_uint(v - __uint_as_float(hi)) + 0x1000u;
}

}  // namespace umma
}  // namespace i2p
