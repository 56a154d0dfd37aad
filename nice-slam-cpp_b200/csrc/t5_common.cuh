// Shared pieces of the tcgen05 decoder kernels (decode_fwd_t5.cu, decode_bwd_t5.cu): mbarrier / tcgen05 wrappers, the UMMA
// shared-memory tile layout of the weights, the three-product fp32-grade issue sequence.
#pragma once
#include "common.cuh"

#ifndef NSB_MBAR_HINT_NS
#define NSB_MBAR_HINT_NS 4000   // suspend-time hint of mbarrier.try_wait in ns (0: the loop spins; measured neutral in time, fewer issued instructions)
#endif

namespace nsb {
namespace t5 {

constexpr int TM = 128;                        // samples per tile (UMMA M)
constexpr int NG = 4;                          // tile groups per CTA
constexpr int GTHREADS = 128;                  // compute threads per group: one per sample
constexpr int CTHREADS = NG * GTHREADS;
constexpr int THREADS = CTHREADS + 32 * NG;    // + one issuer warp per group
// setmaxnreg moves registers inside the CTA's own launch allocation (640 threads x 96 = 61440): 512 x 112 + 128 x 32 = 61440
constexpr int REGS_COMPUTE = 112, REGS_ISSUE = 32;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {   // try_wait sleeps in hardware until the phase flips or a time limit
    asm volatile(
        "{\n\t.reg .pred P1;\n\tWAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n\t"      // suspend-time hint: sleep in hardware instead of spinning
        "@P1 bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(bar), "r"(parity), "r"((unsigned)NSB_MBAR_HINT_NS) : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {   // K-major SWIZZLE_128B, SBO 1024 B, version 1
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {   // kind::f16: A, B = F16 (format 0), D = F32, both K-major
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// D[tmem] (+)= A[tmem] . B[smem]^T
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
// D[tmem] (+)= A[smem] . B[smem]^T
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void group_sync(int grp) { asm volatile("bar.sync %0, %1;" ::"r"(1 + grp), "r"(GTHREADS + 32) : "memory"); }
__device__ __forceinline__ void mma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]),
                   "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]),
                   "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]),
                   "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]),
                   "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// 32 fp32 values of this thread's row -> 32 tensor-memory columns: 16 packed f16x2 words of hi parts, then 16 of lo parts
__device__ __forceinline__ void store_operand(uint32_t taddr, const float (&v)[32]) {
    uint32_t w[32];
#pragma unroll
    for (int i = 0; i < 16; ++i) split_f16(v[2 * i], v[2 * i + 1], w[i], w[16 + i]);
    tmem_st32(taddr, w);
}

// byte offset of the 16-byte chunk `c` (0..3 hi, 4..7 lo) of row r inside a SWIZZLE_128B tile
__device__ __forceinline__ int chunk_off(int r, int c) { return (r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4); }
// one element of a weight tile: row r, logical input index k (0..31) -> hi at half k, lo at half 32 + k
__device__ __forceinline__ void put_w(uint8_t* tile, int r, int k, float v) {
    const __half h = __float2half_rn(v);
    const __half l = __float2half_rn(v - __half2float(h));
    *reinterpret_cast<__half*>(tile + chunk_off(r, k >> 3) + (k & 7) * 2) = h;
    *reinterpret_cast<__half*>(tile + chunk_off(r, 4 + (k >> 3)) + (k & 7) * 2) = l;
}

// fp32-grade product over one 32-input block, A at tensor-memory column `a` (hi 16 columns | lo 16 columns, 8 columns per k16
// step), B a shared-memory tile (descriptor units of 16 B: +2 = one k16 step, +4 = the lo half of the 128-byte row)
__device__ __forceinline__ void issue3(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, bool zero_first) {
#pragma unroll
    for (int k = 0; k < 2; ++k) mma_ts(d, a + 16 + 8 * k, b + 2 * k, idesc, (zero_first && k == 0) ? 0u : 1u);   // a_lo . b_hi (small terms first)
#pragma unroll
    for (int k = 0; k < 2; ++k) mma_ts(d, a + 8 * k, b + 4 + 2 * k, idesc, 1u);                                 // a_hi . b_lo
#pragma unroll
    for (int k = 0; k < 2; ++k) mma_ts(d, a + 8 * k, b + 2 * k, idesc, 1u);                                     // a_hi . b_hi
}

}  // namespace t5
}  // namespace nsb
