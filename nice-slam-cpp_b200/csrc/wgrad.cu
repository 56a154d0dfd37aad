// Split-K weight-gradient kernel for the colour decoder (the only decoder the reference optimises by default:
// fix_fine = True, fix_color = False, Mapper.cpp:292-301).  The backward decoder kernel stashes, per sample, the
// layer inputs X and the layer-output gradients G; every parameter gradient is then a tall-skinny product
//     dW[a][b] = sum_s L[s][a] * R[s][b]
// with the sample index as the contraction dimension.  Each CTA walks a contiguous range of samples, each warp
// owns three groups of (one 16-row A tile) x (four 8-column B tiles) and keeps their accumulators in registers;
// the per-CTA partial sums are added to the flat gradient with fp32 reductions at the end.
#include "decode.cuh"
#include "params.h"

namespace nsb {

constexpr int WG_WARPS = 16;
constexpr int WG_GROUPS_PER_WARP = 3;
constexpr int WG_MAX_GROUPS = WG_WARPS * WG_GROUPS_PER_WARP;   // 48 >= 45

struct WGroup {
    int L, R;          // stash column of row a0 of the A tile / of column b0 of the B tiles
    int dst, ld;       // flat-gradient offset of element (a_lo, 0) and its row stride
    int a_lo, a_hi;    // valid rows of the A tile (relative to L)
    int nb;            // valid columns (relative to R), up to 32
};
struct WGTable { WGroup g[WG_MAX_GROUPS]; int n; };

static WGTable build_table() {
    WGTable T; T.n = 0;
    const DecFlat f = DecFlat::make(32, 4);
    auto add = [&](int L, int R, int dst, int ld, int a_lo, int a_hi, int nb) { T.g[T.n++] = WGroup{L, R, dst, ld, a_lo, a_hi, nb}; };
    auto addmat = [&](int L, int na, int R, int nb, int dst, int ld) {   // dW[a][b], a < na, b < nb
        for (int a0 = 0; a0 < na; a0 += 16)
            for (int b0 = 0; b0 < nb; b0 += 32)
                add(L + a0, R + b0, dst + a0 * ld + b0, ld, 0, (na - a0) < 16 ? (na - a0) : 16, (nb - b0) < 32 ? (nb - b0) : 32);
    };
    addmat(stash::GU + 0, 32, stash::E, EMB, f.W[0], EMB);
    addmat(stash::GU + 32, 32, stash::H + 0, 32, f.W[1], 32);
    addmat(stash::GU + 64, 32, stash::H + 32, 32, f.W[2], 32);
    addmat(stash::GU + 96, 32, stash::E, EMB, f.W[3], EMB + HID);
    addmat(stash::GU + 96, 32, stash::H + 64, 32, f.W[3] + EMB, EMB + HID);
    addmat(stash::GU + 128, 32, stash::H + 96, 32, f.W[4], 32);
    for (int i = 0; i < 5; ++i) addmat(stash::GH + 32 * i, 32, stash::Cc, 32, f.Fc[i], 32);
    addmat(stash::GO, 4, stash::H + 128, 32, f.Wo, 32);
    addmat(stash::Pp, 3, stash::GE, EMB, f.B, EMB);                       // dB[d][f] = sum p_d * g_e cos
    for (int i = 0; i < 5; ++i) add(stash::Pp, stash::GU + 32 * i, f.b[i], 0, 3, 4, 32);    // bias = column sums (Pp[3] == 1)
    for (int i = 0; i < 5; ++i) add(stash::Pp, stash::GH + 32 * i, f.bc[i], 0, 3, 4, 32);
    add(stash::Pp, stash::GO, f.bo, 0, 3, 4, 4);
    return T;
}

__constant__ WGTable c_wg;

constexpr int WG_STAGES = 3;                       // cp.async ring depth (k-steps in flight)
constexpr int WG_KROWS = 16;                       // samples per k-step (MMA k = 16)
constexpr int WG_STAGE_FLOATS = WG_KROWS * stash::W + 32;   // +32: the last B tile of a row may read past it (values unused)

__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gmem_src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

// The CTA streams its sample range through a shared-memory ring (one stage = the 16 stash rows of a k-step,
// 45.6 KB, fetched once with 16-byte cp.async), so every stash element crosses L2/HBM exactly once and the MMA
// fragments come from conflict-free LDS.64 / LDS.128 (row stride 712 = 8 mod 32 floats).  k slots (2t, 2t+1, 2t+8, 2t+9)
// of lane t <-> ring rows (t, t+4, t+8, t+12): any bijection works as long as both operands use it.
template <bool P3>
__global__ void __launch_bounds__(WG_WARPS * 32) k_wgrad(const float* __restrict__ st, const uint8_t* __restrict__ valid,
                                                         int P, int S, float* __restrict__ dflat) {
    extern __shared__ __align__(128) float ring[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    float acc[WG_GROUPS_PER_WARP][4][4];
#pragma unroll
    for (int q = 0; q < WG_GROUPS_PER_WARP; ++q)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[q][j][0] = acc[q][j][1] = acc[q][j][2] = acc[q][j][3] = 0.0f;
    int Lc[WG_GROUPS_PER_WARP], Rc[WG_GROUPS_PER_WARP]; bool on[WG_GROUPS_PER_WARP];
#pragma unroll
    for (int q = 0; q < WG_GROUPS_PER_WARP; ++q) {
        const int gi = warp * WG_GROUPS_PER_WARP + q;
        on[q] = gi < c_wg.n;
        Lc[q] = on[q] ? c_wg.g[gi].L + 2 * g : 0;      // A rows g / g+8 <-> a0 + 2g, a0 + 2g + 1
        Rc[q] = on[q] ? c_wg.g[gi].R + 4 * g : 0;      // B col g of tile j <-> b0 + 4g + j
    }
    const int nks = P / WG_KROWS;
    const int per = (nks + gridDim.x - 1) / gridDim.x;
    const int k_lo = blockIdx.x * per, k_hi = min(nks, k_lo + per);
    const int n_my = max(0, k_hi - k_lo);
    constexpr int CHUNKS = WG_KROWS * stash::W / 4;    // 16-byte chunks per stage
    auto issue = [&](int i) {                          // fetch k-step k_lo + i into ring slot i % WG_STAGES
        if (i < n_my) {
            const float* src = st + (size_t)(k_lo + i) * WG_KROWS * stash::W;
            float* dst = ring + (i % WG_STAGES) * WG_STAGE_FLOATS;
            for (int c = threadIdx.x; c < CHUNKS; c += blockDim.x) cp_async16(dst + 4 * c, src + 4 * c);
        }
        cp_async_commit();
    };
#pragma unroll
    for (int i = 0; i < WG_STAGES - 1; ++i) issue(i);
    for (int i = 0; i < n_my; ++i) {
        cp_async_wait<WG_STAGES - 2>();                // stage i has landed (for this thread's copies) ...
        __syncthreads();                               // ... and for everybody's; also: slot (i-1) % STAGES is free again
        issue(i + WG_STAGES - 1);
        const int s0 = (k_lo + i) * WG_KROWS;
        if (valid && !valid[s0 / S]) continue;         // rows of rays dropped by the inside filter were never written
        const float* r0 = ring + (i % WG_STAGES) * WG_STAGE_FLOATS + t * stash::W;
#pragma unroll
        for (int q = 0; q < WG_GROUPS_PER_WARP; ++q) {
            if (!on[q]) continue;
            float2 av[4]; float4 bv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                av[u] = *reinterpret_cast<const float2*>(r0 + 4 * u * stash::W + Lc[q]);
                bv[u] = *reinterpret_cast<const float4*>(r0 + 4 * u * stash::W + Rc[q]);
            }
            AFrag<P3> a;
            a.set(av[0].x, av[1].x, av[2].x, av[3].x, av[0].y, av[1].y, av[2].y, av[3].y);
            mma_acc<P3>(acc[q][0], a, bv[0].x, bv[1].x, bv[2].x, bv[3].x);
            mma_acc<P3>(acc[q][1], a, bv[0].y, bv[1].y, bv[2].y, bv[3].y);
            mma_acc<P3>(acc[q][2], a, bv[0].z, bv[1].z, bv[2].z, bv[3].z);
            mma_acc<P3>(acc[q][3], a, bv[0].w, bv[1].w, bv[2].w, bv[3].w);
        }
    }
    cp_async_wait<0>();
#pragma unroll
    for (int q = 0; q < WG_GROUPS_PER_WARP; ++q) {
        if (!on[q]) continue;
        const WGroup G = c_wg.g[warp * WG_GROUPS_PER_WARP + q];
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int a = 2 * g + (e >> 1), b = 4 * (2 * t + (e & 1)) + j;
                if (a >= G.a_lo && a < G.a_hi && b < G.nb) atomicAdd(dflat + G.dst + (a - G.a_lo) * G.ld + b, acc[q][j][e]);
            }
    }
}

cudaError_t launch_wgrad(const float* stash_buf, const uint8_t* valid, int P, int S, float* dflat, int precision, int grid, cudaStream_t st) {
    static bool init = false;
    if (!init) {
        const WGTable T = build_table();
        cudaError_t e = cudaMemcpyToSymbol(c_wg, &T, sizeof(T));
        if (e != cudaSuccess) return e;
    }
    const size_t smem = sizeof(float) * WG_STAGES * WG_STAGE_FLOATS;
    if (!init) {
        cudaError_t e = cudaFuncSetAttribute(k_wgrad<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_wgrad<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        init = true;
    }
    if (precision == 0) k_wgrad<true><<<grid, WG_WARPS * 32, smem, st>>>(stash_buf, valid, P, S, dflat);
    else k_wgrad<false><<<grid, WG_WARPS * 32, smem, st>>>(stash_buf, valid, P, S, dflat);
    return cudaGetLastError();
}

}  // namespace nsb
