// Multi-GPU optimiser step as ONE kernel over NVLink peer memory: reduce-scatter of the gradient arena (each rank sums its
// slice of all ranks' gradients with P2P loads), Adam on the owned slice (identical arithmetic to k_adam), all-gather of the
// updated parameters (P2P stores into every rank's parameter arena), bracketed by two flag barriers that live in peer memory.
// Replaces  ncclAllReduce(grad) + k_adam  of the mapping iteration (C1 + K12/K13 of SURVEY 2.1): the exchange moves 2 x (W-1)/W
// of the arena per rank instead of an all-reduce's traffic, the Adam pass shrinks to 1/W of the arena, and there is no
// launch or protocol latency of a separate collective.
#pragma once
#include "misc_kernels.cuh"

namespace nsb {

constexpr int P2P_MAX_WORLD = 8;

struct P2PParams {
    AdamParams A;                          // segments, betas, step sizes (param / grad / m / v = this rank's arenas)
    const float* peer_grad[P2P_MAX_WORLD]; // every rank's gradient arena (index = rank; own entry = local pointer)
    float* peer_param[P2P_MAX_WORLD];      // every rank's parameter arena
    uint32_t* peer_flags[P2P_MAX_WORLD];   // every rank's flag block: [0..7] ready epochs, [8..15] done epochs, [16] CTA counter
    int rank, world;
    int lo4, hi4;                          // the exchanged range of the arena in float4 units (ownership: see the kernel)
    int loss4;                             // float4 index of the loss slot (summed, written to param arenas, no Adam), or -1
    uint32_t epoch;
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) { uint32_t v; asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ float4 ld_peer(const float* p) {   // remote line: bypass L1 (it may hold last iteration's value)
    float4 v; asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p)); return v;
}

__global__ void __launch_bounds__(256) k_reduce_adam(P2PParams P) {
    uint32_t* my_flags = P.peer_flags[P.rank];
    // ---- barrier 1: every rank's backward has finished (this kernel is behind it in each rank's stream)
    if (blockIdx.x == 0 && threadIdx.x < P.world) st_release_sys(P.peer_flags[threadIdx.x] + P.rank, P.epoch);
    if (threadIdx.x < P.world) { while (ld_acquire_sys(my_flags + threadIdx.x) < P.epoch) { } }
    __syncthreads();

    // Ownership is static and interleaved: block b of 256 float4 (4 KB) of the arena belongs to rank b % world, whatever range
    // an iteration exchanges -- the owner alone keeps Adam's m and v of a block, so it must never change, and any exchanged
    // range (the short geometry-stage prefix or the whole arena) is spread evenly over the ranks.
    constexpr int BLK4 = 256;
    const int b_lo = P.lo4 / BLK4;
    const int first = b_lo + ((P.rank - b_lo % P.world) + P.world) % P.world;      // first block >= b_lo owned by this rank
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; ; j += gridDim.x * blockDim.x) {
        const int i4 = (first + (j / BLK4) * P.world) * BLK4 + (j % BLK4);
        if (i4 >= P.hi4) break;
        if (i4 < P.lo4) continue;
        // segment of this element (compile-time indices only, as in k_adam)
        AdamSegment sg = P.A.seg[0];
#pragma unroll
        for (int k = 1; k < ADAM_MAX_SEG; ++k)
            if (k < P.A.n_seg && 4 * i4 >= P.A.seg[k].begin && 4 * i4 < P.A.seg[k].end) sg = P.A.seg[k];
        const bool in_seg = 4 * i4 >= sg.begin && 4 * i4 < sg.end;
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int w = 0; w < P2P_MAX_WORLD; ++w) {
            if (w < P.world) { const float4 x = ld_peer(P.peer_grad[w] + 4 * (size_t)i4); g.x += x.x; g.y += x.y; g.z += x.z; g.w += x.w; }
        }
        float4 p;
        bool write = false;
        if (i4 == P.loss4) { p = g; write = true; }
        else if (in_seg && sg.active && !(sg.mask && !sg.mask[(i4 * 4 - sg.begin) / CDIM])) {
            float4 m = reinterpret_cast<const float4*>(P.A.m)[i4], v = reinterpret_cast<const float4*>(P.A.v)[i4];
            p = reinterpret_cast<const float4*>(P.A.param)[i4];
            float* gg = reinterpret_cast<float*>(&g); float* mm = reinterpret_cast<float*>(&m);
            float* vv = reinterpret_cast<float*>(&v); float* pp = reinterpret_cast<float*>(&p);
            const float nstep = -sg.step;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float gk = gg[k] * P.A.grad_scale;
                mm[k] = __fadd_rn(__fmul_rn(mm[k], P.A.beta1), __fmul_rn(P.A.om_beta1, gk));
                vv[k] = __fadd_rn(__fmul_rn(vv[k], P.A.beta2), __fmul_rn(__fmul_rn(P.A.om_beta2, gk), gk));
                const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(vv[k]), P.A.bc2_sqrt), P.A.eps);
                pp[k] = __fadd_rn(pp[k], __fdiv_rn(__fmul_rn(nstep, mm[k]), denom));
            }
            reinterpret_cast<float4*>(P.A.m)[i4] = m; reinterpret_cast<float4*>(P.A.v)[i4] = v;
            write = true;
        }
        if (write) {
#pragma unroll
            for (int w = 0; w < P2P_MAX_WORLD; ++w)
                if (w < P.world) __stcg(reinterpret_cast<float4*>(P.peer_param[w]) + i4, p);
        }
    }
    // ---- barrier 2: this rank's reads of the peers' gradients and writes into their parameters are complete; the last CTA
    // tells the peers and then waits until every peer has said the same, so that when this kernel retires (a) all parameters
    // of this rank are up to date and (b) nobody reads this rank's gradient arena any more (it is cleared right after).
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t done = atomicAdd(my_flags + 16, 1u) + 1u;
        if (done == gridDim.x) {
            my_flags[16] = 0;
            __threadfence_system();
            for (int w = 0; w < P.world; ++w) st_release_sys(P.peer_flags[w] + 8 + P.rank, P.epoch);
            for (int w = 0; w < P.world; ++w) { while (ld_acquire_sys(my_flags + 8 + w) < P.epoch) { } }
        }
    }
}

}  // namespace nsb
