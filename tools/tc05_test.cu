// Stand-alone validation of the tcgen05 building blocks used by the tensor-memory decoder kernels:
//   * shared-memory operand layout (K-major, SWIZZLE_128B) + 64-bit matrix descriptor + 32-bit instruction descriptor
//   * tcgen05.mma kind::tf32, A from shared memory (SS) and A from tensor memory (TS), N = 32 and N = 64
//   * tcgen05.alloc / commit -> mbarrier / ld 32x32b.x32 / st 32x32b.x32 and the fences between them
// Prints max |err| of each variant against a host reference.  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/tc05_test tools/tc05_test.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred P1;\n\tWAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
    // K-major, SWIZZLE_128B: start address >> 4, LBO = 1 (unused), SBO = 1024 B >> 4, version = 1 (Blackwell), layout type 2
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N) {
    // c_format F32 (1) @4, a_format TF32 (2) @7, b_format TF32 (2) @10, K-major A and B, N>>3 @17, M>>4 @24
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]),
                   "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]),
                   "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float* v) {
    const uint32_t* r = reinterpret_cast<const uint32_t*>(v);
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]),
                   "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]),
                   "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
// float offset of element (row r, k) of a [rows][32] fp32 tile in the K-major SWIZZLE_128B layout (8-row atoms of 1 KB)
__host__ __device__ inline int sw128(int r, int k) { return (r >> 3) * 256 + (r & 7) * 32 + (((k >> 2) ^ (r & 7)) << 2) + (k & 3); }

__global__ void __launch_bounds__(128) k_test(const float* __restrict__ A, const float* __restrict__ B32, const float* __restrict__ B64,
                                              float* out_ss, float* out_ts, float* out_n64) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    float* sA = reinterpret_cast<float*>(smem_raw);            // [128][32]  16 KB
    float* sB = sA + 128 * 32;                                   // [32][32]    4 KB
    float* sB64 = sB + 32 * 32;                                  // [64][32]    8 KB
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(256u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;"); }
    for (int i = tid; i < 128 * 32; i += 128) sA[sw128(i / 32, i % 32)] = A[i];
    for (int i = tid; i < 32 * 32; i += 128) sB[sw128(i / 32, i % 32)] = B32[i];
    for (int i = tid; i < 64 * 32; i += 128) sB64[sw128(i / 32, i % 32)] = B64[i];
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy smem writes -> visible to the tensor core
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    // A operand also into tensor memory (columns 32..63): one row per thread
    {
        float row[32];
        for (int k = 0; k < 32; ++k) row[k] = A[tid * 32 + k];
        tmem_st32(tmem + lane_base + 32, row);
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint64_t da = make_desc_sw128(smem_u32(sA)), db = make_desc_sw128(smem_u32(sB)), db64 = make_desc_sw128(smem_u32(sB64));
        for (int k = 0; k < 4; ++k) mma_ss(tmem + 0, da + 2 * k, db + 2 * k, make_idesc_tf32(128, 32), k > 0);           // D0: cols 0..31
        for (int k = 0; k < 4; ++k) mma_ts(tmem + 64, tmem + 32 + 8 * k, db + 2 * k, make_idesc_tf32(128, 32), k > 0);  // D1: cols 64..95
        for (int k = 0; k < 4; ++k) mma_ss(tmem + 128, da + 2 * k, db64 + 2 * k, make_idesc_tf32(128, 64), k > 0);       // D2: cols 128..191
        mma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    float v[32];
    tmem_ld32(tmem + lane_base + 0, v);
    for (int k = 0; k < 32; ++k) out_ss[tid * 32 + k] = v[k];
    tmem_ld32(tmem + lane_base + 64, v);
    for (int k = 0; k < 32; ++k) out_ts[tid * 32 + k] = v[k];
    tmem_ld32(tmem + lane_base + 128, v);
    for (int k = 0; k < 32; ++k) out_n64[tid * 64 + k] = v[k];
    tmem_ld32(tmem + lane_base + 160, v);
    for (int k = 0; k < 32; ++k) out_n64[tid * 64 + 32 + k] = v[k];
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u));
}

static float tf32_round(float x) { uint32_t u; memcpy(&u, &x, 4); u = (u + 0x1000u) & 0xffffe000u; float r; memcpy(&r, &u, 4); return r; }
static float tf32_trunc(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xffffe000u; float r; memcpy(&r, &u, 4); return r; }

int main() {
    std::vector<float> A(128 * 32), B(32 * 32), B64(64 * 32);
    srand(1);
    auto rnd = [] { return (float)rand() / RAND_MAX * 2.f - 1.f; };
    for (int pass = 0; pass < 2; ++pass) {      // pass 0: inputs already tf32-exact; pass 1: full fp32 inputs (learn trunc vs round)
        for (auto& x : A) x = pass == 0 ? tf32_round(rnd()) : rnd();
        for (auto& x : B) x = pass == 0 ? tf32_round(rnd()) : rnd();
        for (auto& x : B64) x = pass == 0 ? tf32_round(rnd()) : rnd();
        float *dA, *dB, *dB64, *o0, *o1, *o2;
        cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dB64, B64.size() * 4);
        cudaMalloc(&o0, 128 * 32 * 4); cudaMalloc(&o1, 128 * 32 * 4); cudaMalloc(&o2, 128 * 64 * 4);
        cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
        cudaMemcpy(dB64, B64.data(), B64.size() * 4, cudaMemcpyHostToDevice);
        cudaMemset(o0, 0, 128 * 32 * 4); cudaMemset(o1, 0, 128 * 32 * 4); cudaMemset(o2, 0, 128 * 64 * 4);
        const int smem = (128 * 32 + 32 * 32 + 64 * 32) * 4 + 1024;
        cudaFuncSetAttribute(k_test, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        k_test<<<1, 128, smem>>>(dA, dB, dB64, o0, o1, o2);
        cudaError_t e = cudaDeviceSynchronize();
        std::vector<float> h0(128 * 32), h1(128 * 32), h2(128 * 64);
        cudaMemcpy(h0.data(), o0, h0.size() * 4, cudaMemcpyDeviceToHost); cudaMemcpy(h1.data(), o1, h1.size() * 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(h2.data(), o2, h2.size() * 4, cudaMemcpyDeviceToHost);
        double e_ss = 0, e_ts = 0, e_64 = 0, e_ss_tr = 0, e_ss_rn = 0;
        for (int m = 0; m < 128; ++m) {
            for (int n = 0; n < 64; ++n) {
                double ref = 0, ref_tr = 0, ref_rn = 0;
                const float* b = n < 32 ? &B[n * 32] : nullptr;
                for (int k = 0; k < 32; ++k) {
                    ref += (double)A[m * 32 + k] * B64[n * 32 + k];
                }
                e_64 = fmax(e_64, fabs(ref - h2[m * 64 + n]));
                if (b) {
                    ref = 0;
                    for (int k = 0; k < 32; ++k) {
                        ref += (double)A[m * 32 + k] * b[k];
                        ref_tr += (double)tf32_trunc(A[m * 32 + k]) * tf32_trunc(b[k]);
                        ref_rn += (double)tf32_round(A[m * 32 + k]) * tf32_round(b[k]);
                    }
                    e_ss = fmax(e_ss, fabs(ref - h0[m * 32 + n])); e_ts = fmax(e_ts, fabs(ref - h1[m * 32 + n]));
                    e_ss_tr = fmax(e_ss_tr, fabs(ref_tr - h0[m * 32 + n])); e_ss_rn = fmax(e_ss_rn, fabs(ref_rn - h0[m * 32 + n]));
                }
            }
        }
        printf("{\"pass\": %d, \"cuda\": \"%s\", \"err_ss\": %.3e, \"err_ts\": %.3e, \"err_n64\": %.3e, \"err_ss_vs_trunc\": %.3e, \"err_ss_vs_round\": %.3e, \"sample\": [%.5f, %.5f, %.5f]}\n",
               pass, cudaGetErrorString(e), e_ss, e_ts, e_64, e_ss_tr, e_ss_rn, h0[0], h1[0], h2[0]);
        if (e != cudaSuccess) return 1;
    }
    return 0;
}
