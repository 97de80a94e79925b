// tcgen05.cuh -- thin inline-PTX layer over the 5th-generation tensor cores (tcgen05.mma, TMEM, mbarrier) used by
// the fused MLP.  No CUTLASS: descriptors are built by hand for the one shared-memory layout the MLP uses.
//
// Shared-memory operand layout ("row-panel", no swizzle): a tile with R rows (multiple of 8) and F columns
// (multiple of 8) of 16-bit elements is stored as F/8 panels of R*16 bytes; inside a panel row r holds its 8
// consecutive columns in 16 bytes:
//        addr(r, c) = (c / 8) * (R * 16) + r * 16 + (c % 8) * 2
// Every 8x16-byte block is a UMMA "core matrix".  The same bytes can be consumed
//   * K-major  (MN = row,    K = column): SBO (MN-direction core-matrix stride) = 128 B,  LBO (K-direction) = R*16 B
//   * MN-major (MN = column, K = row)   : SBO = R*16 B,                                   LBO = 128 B
// so one activation tile feeds both  Y = A * W^T  (contract over columns) and  dW^T = A^T * dZ  (contract over
// rows) without a transposed copy.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

namespace ngp {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// 64-bit shared-memory matrix descriptor (SWIZZLE_NONE, Blackwell version field = 1)
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

// 32-bit instruction descriptor for kind::f16, fp16 inputs, fp32 accumulate
__device__ __forceinline__ uint32_t instr_desc(uint32_t M, uint32_t N, bool a_mn_major, bool b_mn_major) {
    return (1u << 4) | ((a_mn_major ? 1u : 0u) << 15) | ((b_mn_major ? 1u : 0u) << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread
__device__ __forceinline__ void mma_f16_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

// arrives on the mbarrier when all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void mma_commit(uint32_t mbar_saddr) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar_saddr) : "memory");
}

__device__ __forceinline__ void mbar_init(uint32_t mbar_saddr, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar_saddr), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive(uint32_t mbar_saddr) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(mbar_saddr) : "memory");
}
// expect `bytes` of async (bulk copy) traffic and arrive once
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t mbar_saddr, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar_saddr), "r"(bytes) : "memory");
}
// named barrier among `count` threads (count a multiple of 32); id 0 is __syncthreads
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t count) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
// bulk async copy global -> shared (the TMA engine, no tensor map: 1-D, 16-byte granular), completion on an mbarrier
__device__ __forceinline__ void bulk_g2s(uint32_t dst_saddr, const void* src, uint32_t bytes, uint32_t mbar_saddr) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_saddr), "l"(src), "r"(bytes), "r"(mbar_saddr) : "memory");
}

// bulk async copy shared -> global (TMA engine, 1-D); completion tracked by the issuing thread's bulk async-group
__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src_saddr, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src_saddr), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
// all of this thread's bulk stores have finished READING shared memory (the source may be overwritten)
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// all of this thread's bulk stores are complete
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// Bounded wait: a mis-programmed pipeline traps instead of hanging the GPU.  try_wait carries a suspend-time hint so that a
// waiting warp sleeps in hardware instead of burning issue slots that the working warps of the CTA need (in the
// warp-specialised kernels up to a quarter of all issued instructions were wait loops before this).
__device__ __forceinline__ void mbar_wait(uint32_t mbar_saddr, uint32_t parity) {
    uint32_t done = 0;
    for (uint32_t spin = 0; !done; ++spin) {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}\n" : "=r"(done) : "r"(mbar_saddr), "r"(parity), "r"(100000u) : "memory");
        if (spin > (1u << 22)) __trap();
    }
}

__device__ __forceinline__ void tmem_alloc(uint32_t dst_saddr, uint32_t ncols) {  // whole warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_saddr), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // whole warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// generic-proxy smem writes -> visible to the async proxy (tensor core reads operands through it)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// 16 consecutive fp32 columns of this thread's TMEM lane (lane = 32 * (warp % 4) + laneid)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int i = 0; i < 16; i++) v[i] = __uint_as_float(r[i]);
}

// 2 consecutive fp32 columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld2(uint32_t taddr, float (&v)[2]) {
    uint32_t r0, r1;
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(r0), "=r"(r1) : "r"(taddr) : "memory");
    v[0] = __uint_as_float(r0); v[1] = __uint_as_float(r1);
}

// 8 consecutive fp32 columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void cp_async16(uint32_t dst_saddr, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_saddr), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

}  // namespace tc
}  // namespace ngp
