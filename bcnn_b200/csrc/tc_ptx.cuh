// tc_ptx.cuh -- thin inline-PTX wrappers for the Blackwell tensor-core path (sm_100a):
// mbarrier, bulk async copy (TMA engine), proxy fences, TMEM allocation, tcgen05.mma /
// commit / ld, and the UMMA shared-memory / instruction descriptors.
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

namespace b200 {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.b32 %0, 1, 0, p;\n"
            "}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
}
// One thread of a converged warp (the same one every time): the issuer of tcgen05.mma / commit.
// ptxas keeps warp-uniform operands of the guarded instructions in uniform registers, which an
// `if (lane == 0)` branch does not allow (it falls back to an ELECT / R2UR.BROADCAST loop per MMA).
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "elect.sync _|p, 0xffffffff;\n"
        "selp.b32 %0, 1, 0, p;\n"
        "}\n" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void bulk_copy_g2s(uint32_t dst_smem, const void *src, uint32_t bytes,
                                              uint32_t bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        :: "r"(dst_smem), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// 16-byte store to shared memory by 32-bit shared address (a store through a generic pointer compiles
// to ST.E with 64-bit address arithmetic)
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" :: "r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 :: "r"(dst_smem), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(taddr), "r"(ncols)
                 : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, BF16 inputs, FP32 accumulate.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                          uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" :: "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// Arrive on an mbarrier once every previously issued tcgen05.mma of this thread completed.
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                 :: "r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,"
        "%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
          "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
          "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]),
          "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]),
          "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Fragment-shaped TMEM loads of a warp's 32 lanes x 8N columns as two 16-lane halves (lo: lanes of
// taddr, hi: 16 lanes further), both in flight before the one wait. tcgen05.ld.16x256b.xN gives thread t,
// for every 8-column block b,
//   v[4b + {0,1}] = (lane t/4,     columns 8b + 2(t%4) + {0,1})
//   v[4b + {2,3}] = (lane t/4 + 8, same columns)
// (the m16n8 accumulator fragment).
__device__ __forceinline__ void tmem_ld_frag64(uint32_t taddr, uint32_t (&lo)[32], uint32_t (&hi)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.16x256b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%64];\n"
        "tcgen05.ld.sync.aligned.16x256b.x8.b32 {%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63}, [%65];\n"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(lo[0]), "=r"(lo[1]), "=r"(lo[2]), "=r"(lo[3]), "=r"(lo[4]), "=r"(lo[5]), "=r"(lo[6]), "=r"(lo[7]), "=r"(lo[8]), "=r"(lo[9]), "=r"(lo[10]), "=r"(lo[11]), "=r"(lo[12]), "=r"(lo[13]), "=r"(lo[14]), "=r"(lo[15]), "=r"(lo[16]), "=r"(lo[17]), "=r"(lo[18]), "=r"(lo[19]), "=r"(lo[20]), "=r"(lo[21]), "=r"(lo[22]), "=r"(lo[23]), "=r"(lo[24]), "=r"(lo[25]), "=r"(lo[26]), "=r"(lo[27]), "=r"(lo[28]), "=r"(lo[29]), "=r"(lo[30]), "=r"(lo[31]),
          "=r"(hi[0]), "=r"(hi[1]), "=r"(hi[2]), "=r"(hi[3]), "=r"(hi[4]), "=r"(hi[5]), "=r"(hi[6]), "=r"(hi[7]), "=r"(hi[8]), "=r"(hi[9]), "=r"(hi[10]), "=r"(hi[11]), "=r"(hi[12]), "=r"(hi[13]), "=r"(hi[14]), "=r"(hi[15]), "=r"(hi[16]), "=r"(hi[17]), "=r"(hi[18]), "=r"(hi[19]), "=r"(hi[20]), "=r"(hi[21]), "=r"(hi[22]), "=r"(hi[23]), "=r"(hi[24]), "=r"(hi[25]), "=r"(hi[26]), "=r"(hi[27]), "=r"(hi[28]), "=r"(hi[29]), "=r"(hi[30]), "=r"(hi[31])
        : "r"(taddr), "r"(taddr + (16u << 16)) : "memory");
}
__device__ __forceinline__ void tmem_ld_frag32(uint32_t taddr, uint32_t (&lo)[16], uint32_t (&hi)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%32];\n"
        "tcgen05.ld.sync.aligned.16x256b.x4.b32 {%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%33];\n"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(lo[0]), "=r"(lo[1]), "=r"(lo[2]), "=r"(lo[3]), "=r"(lo[4]), "=r"(lo[5]), "=r"(lo[6]), "=r"(lo[7]), "=r"(lo[8]), "=r"(lo[9]), "=r"(lo[10]), "=r"(lo[11]), "=r"(lo[12]), "=r"(lo[13]), "=r"(lo[14]), "=r"(lo[15]),
          "=r"(hi[0]), "=r"(hi[1]), "=r"(hi[2]), "=r"(hi[3]), "=r"(hi[4]), "=r"(hi[5]), "=r"(hi[6]), "=r"(hi[7]), "=r"(hi[8]), "=r"(hi[9]), "=r"(hi[10]), "=r"(hi[11]), "=r"(hi[12]), "=r"(hi[13]), "=r"(hi[14]), "=r"(hi[15])
        : "r"(taddr), "r"(taddr + (16u << 16)) : "memory");
}

// K-major, 128-byte-swizzled shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
// start address >> 4 in [0,14), LBO (unused for swizzled K-major) = 1 in [16,30),
// SBO = 1024 B (8 rows x 128 B) >> 4 in [32,46), version = 1 in [46,48), layout type
// SWIZZLE_128B = 2 in [61,64).
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): D = F32 (1 << 4), A = B = BF16
// (1 << 7, 1 << 10), both K-major, N >> 3 at [17,23), M >> 4 at [24,29).
__device__ __forceinline__ uint32_t make_idesc(int m, int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) |
           ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}


}  // namespace tc
}  // namespace b200
