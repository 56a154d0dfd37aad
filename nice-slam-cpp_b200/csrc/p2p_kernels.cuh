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
    uint32_t* peer_flags[P2P_MAX_WORLD];   // every rank's flag block: [0..7] ready epochs, [8..15] done epochs, [16] CTA counter,
                                           // [17] this rank's own epoch (number of exchanges done: the kernel reads its epoch from
                                           // here so that a captured graph can be replayed), [18] barrier time-out flag,
                                           // [20..27] globaltimer durations in ns as 64-bit values: last barrier-1 wait, last kernel
                                           // total, sum of the waits, sum of the totals; [28] number of exchanges summed; [30..31] sum of the exchanged range
                                           // sizes in bytes (64-bit)
    int rank, world;
    int lo4, hi4;                          // the exchanged range of the arena in float4 units (ownership: see the kernel)
    int loss4;                             // float4 index of the loss slot (summed, written to param arenas, no Adam), or -1
    float* stats_base;                     // statistics ring: the summed loss goes to slot[3] of the current step
    unsigned long long timeout_ns;         // a barrier that waits longer than this gives up and raises flag [18] (0: wait for ever)
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) { uint32_t v; asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ float4 ld_peer(const float* p) {   // remote line: bypass L1 (it may hold last iteration's value)
    float4 v; asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p)); return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
// Bounded spin on a flag another rank writes: a rank that returned an error before launching its kernel must not hang its peers
// for ever.  Returns false on time-out.
__device__ __forceinline__ bool wait_flag(const uint32_t* p, uint32_t epoch, unsigned long long t0, unsigned long long timeout_ns) {
    unsigned spins = 0;
    while ((int)(ld_acquire_sys(p) - epoch) < 0) {
        if (timeout_ns && (++spins & 1023u) == 0 && globaltimer_ns() - t0 > timeout_ns) return false;
    }
    return true;
}

__global__ void __launch_bounds__(256) k_reduce_adam(P2PParams P) {
    uint32_t* my_flags = P.peer_flags[P.rank];
    const uint32_t epoch = my_flags[17] + 1u;          // local, advanced by the last CTA below
    const unsigned long long t_start = globaltimer_ns();
    double bc1; float bc2s;
    adam_bias(P.A, bc1, bc2s);
    const int slot = iter_slot(P.A.it);
    // ---- barrier 1: every rank's backward has finished (this kernel is behind it in each rank's stream)
    if (blockIdx.x == 0 && threadIdx.x < P.world) st_release_sys(P.peer_flags[threadIdx.x] + P.rank, epoch);
    if (threadIdx.x < P.world) { if (!wait_flag(my_flags + threadIdx.x, epoch, t_start, P.timeout_ns)) my_flags[18] = 1u; }
    __syncthreads();
    const unsigned long long t_wait = globaltimer_ns();

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
            // gradient sum, m and v all zero: the update is the identity on every replica (see all_zero4) -- nothing to compute or send
            if (all_zero4(g) && all_zero4(m) && all_zero4(v)) continue;
            p = reinterpret_cast<const float4*>(P.A.param)[i4];
            float* gg = reinterpret_cast<float*>(&g); float* mm = reinterpret_cast<float*>(&m);
            float* vv = reinterpret_cast<float*>(&v); float* pp = reinterpret_cast<float*>(&p);
            const float nstep = -(float)((double)sg.lr / bc1);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float gk = gg[k] * P.A.grad_scale;
                mm[k] = __fadd_rn(__fmul_rn(mm[k], P.A.beta1), __fmul_rn(P.A.om_beta1, gk));
                vv[k] = __fadd_rn(__fmul_rn(vv[k], P.A.beta2), __fmul_rn(__fmul_rn(P.A.om_beta2, gk), gk));
                const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(vv[k]), bc2s), P.A.eps);
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
            for (int w = 0; w < P.world; ++w) st_release_sys(P.peer_flags[w] + 8 + P.rank, epoch);
            for (int w = 0; w < P.world; ++w) { if (!wait_flag(my_flags + 8 + w, epoch, t_wait, P.timeout_ns)) my_flags[18] = 1u; }
            my_flags[17] = epoch;
            // instrumentation: time spent waiting for the slowest rank's backward (barrier 1) and the whole kernel, in ns
            unsigned long long* ts = reinterpret_cast<unsigned long long*>(my_flags + 20);
            const unsigned long long t_total = globaltimer_ns() - t_start;
            ts[0] = t_wait - t_start; ts[1] = t_total; ts[2] += t_wait - t_start; ts[3] += t_total; my_flags[28] += 1u;
            *reinterpret_cast<unsigned long long*>(my_flags + 30) += 16ull * (unsigned long long)(P.hi4 - P.lo4);
            // every rank's owner has stored the summed loss into this rank's parameter arena before it signalled barrier 2
            if (P.stats_base && P.loss4 >= 0) P.stats_base[4 * slot + 3] = __ldcg(P.A.param + 4 * (size_t)P.loss4);
            if (P.A.it.state) iter_advance(P.A.it, P.stats_base);
        }
    }
}

}  // namespace nsb
