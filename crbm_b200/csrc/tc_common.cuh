// tcgen05 / TMEM / UMMA-descriptor helpers shared by the tensor-memory kernels (scan_tc.cu, train_tc.cu), sm_100a.
#pragma once
#include <cuda_fp16.h>
#include "common.cuh"

namespace crbm {

// ---- PTX wrappers -------------------------------------------------------------------------------------------------
// true in exactly one lane of a converged warp (elect.sync): ptxas knows that code under this predicate runs in one thread, so a
// tcgen05.mma there is one UTCHMMA; under `lane == 0` it wraps every MMA in an elect-and-branch loop over the active lanes
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "@p mov.u32 %0, 1;\n\t}" : "+r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <int COLS_UNUSED = 0>
__device__ __forceinline__ void tmem_alloc(uint32_t *slot_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_smem)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, fp16 inputs, fp32 accumulate; issued by one thread for the CTA
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// The same with the descriptors given by their LOW words only. The issuing thread's instruction stream is the bottleneck of kernels that
// issue many small MMAs (measured: ~140 cycles per MMA when every descriptor is rebuilt from its address: shift, mask, or on the uniform
// datapath, each a dependent ~10-cycle step): the high word of every SWIZZLE_128B descriptor used here is the same constant (stride
// byte offset 1024, version 1, layout 2), and the low word of `address + delta` is the low word of `address` plus `delta >> 4` (the
// 14-bit address field never overflows below 256 KB of shared memory), so a k-step costs one add per operand.
constexpr uint32_t UMMA_DESC_HI_SW128 = (1024u >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint32_t umma_desc_lo(uint32_t smem_addr, uint32_t lbo_bytes = 16u) {
    return ((smem_addr >> 4) & 0x3FFFu) | ((lbo_bytes >> 4) << 16);
}
__device__ __forceinline__ void umma_f16_lo(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %5};\n\t"
        "mov.b64 db, {%2, %5};\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(UMMA_DESC_HI_SW128) : "memory");
}
// arrive on an mbarrier when all tcgen05.mma issued so far by this thread have completed (implies fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// after tmem_ld_wait(): pins the uses of v behind the wait (the compiler sees no data dependence on the wait itself)
__device__ __forceinline__ void tmem_ld_fence8(uint32_t (&v)[8]) {
    asm volatile("" : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]) :: "memory");
}

// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): rows of 128 bytes, 8-row groups
// 1024 bytes apart (stride byte offset), 16-byte chunks XOR-swizzled with the row index inside a group.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);          // start address, bits [0,14)
    d |= (uint64_t)1 << 16;                              // leading byte offset (ignored for swizzled K-major), bits [16,30)
    d |= (uint64_t)(1024 >> 4) << 32;                    // stride byte offset, bits [32,46)
    d |= (uint64_t)1 << 46;                              // descriptor version (Blackwell), bits [46,48)
    d |= (uint64_t)2 << 61;                              // layout type SWIZZLE_128B, bits [61,64)
    return d;
}
// byte offset of 16-byte chunk `chunk` (0..7) of row `r` inside one swizzle atom column (rows x 128 bytes)
__device__ __forceinline__ uint32_t sw128_off(int r, int chunk) {
    return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + (((chunk ^ r) & 7) << 4));
}

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// MN-major SWIZZLE_128B descriptor (cute::UMMA make_umma_desc<Major::MN>: canonical layout, in 16-byte units,
// Swizzle<3,4,3> o ((8,n),(8,k)) : ((1,LBO),(8,SBO))): one K index (row) is 128 bytes = 64 contiguous MN elements, 8 rows form
// a 1024-byte swizzle atom, the next 8 K rows are `sbo_bytes` further, the next 64 MN elements `lbo_bytes` further.
// It reads a tile stored [K rows][MN contiguous] — e.g. the one-hot tile [position][c] used as the operand whose
// contraction index is the position.
__device__ __forceinline__ uint64_t umma_desc_sw128_mn(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

}  // namespace crbm
