// Layout conversion, fused Adam over the parameter arena, small utilities.
#pragma once
#include "common.cuh"

namespace nsb {

// (1,C,Z,Y,X) channel-first  <->  [Z][Y][X][C] channel-last
__global__ void k_ncdhw_to_cl(const float* __restrict__ src, float* __restrict__ dst, int C, int nvox) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;   // i = vox * C + c
    if (i < C * nvox) { const int v = i / C, c = i % C; dst[i] = src[(size_t)c * nvox + v]; }
}
__global__ void k_cl_to_ncdhw(const float* __restrict__ src, float* __restrict__ dst, int C, int nvox) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;   // i = c * nvox + vox
    if (i < C * nvox) { const int c = i / nvox, v = i % nvox; dst[i] = src[(size_t)v * C + c]; }
}

__global__ void k_iota(int64_t* __restrict__ dst, int64_t base, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = base + i;
}

// One contiguous run of the parameter arena that shares a learning rate (an Adam "param group" of
// Mapper.cpp:330: decoders, coarse, middle, fine, color grids, camera tensors).
struct AdamSegment {
    int begin, end;          // float offsets into the arena, multiples of 4
    float lr;                // learning rate of the group; the kernel forms libtorch's step_size = lr / (1 - beta1^t) in double
    const uint8_t* mask;     // per-voxel mask (grids only; element i -> voxel (i - begin) / 32), or nullptr
    int active;              // 0: parameter is not in the optimiser (fix_fine / fix_color): only its gradient is cleared
};
constexpr int ADAM_MAX_SEG = 8;
struct AdamParams {
    float* param; float* grad; float* m; float* v;
    AdamSegment seg[ADAM_MAX_SEG];
    int cum4[ADAM_MAX_SEG + 1];   // prefix sums of the segment lengths in float4 units: thread i works on segment s with cum4[s] <= i < cum4[s+1]
    int n_seg;
    float beta1, beta2, om_beta1, om_beta2, eps;   // om_* = (float)(1.0 - beta) as libtorch passes them
    double bc1;              // 1 - beta1^t   (host, double)                    } used when there is no iteration state;
    float bc2_sqrt;          // (float)sqrt(1 - beta2^t), double arithmetic on the host } the mapping loop reads row state[0] of the tables
    const double* bc1_tab; const float* bc2s_tab; int tab_n;   // per-step bias corrections, t = 1 .. tab_n (row t - 1)
    IterRef it;              // mapping loop: step = it.state[0] (completed iterations), advanced by the last block of this kernel
    float grad_scale;        // 1 (single GPU) -- kept for mean-style reductions
    float* loss_dst;         // optional: receives the loss scalar that rides at float4 slot loss_idx4 of the gradient arena
                             // (with iteration state: base of the statistics ring, the loss goes to slot[3])
    int loss_idx4;
};

// Bias corrections of the current step (libtorch computes them in double from the step count).
__device__ __forceinline__ void adam_bias(const AdamParams& P, double& bc1, float& bc2s) {
    bc1 = P.bc1; bc2s = P.bc2_sqrt;
    if (P.it.state) { const int t = min(P.it.state[0], P.tab_n - 1); bc1 = P.bc1_tab[t]; bc2s = P.bc2s_tab[t]; }
}
// End of a joint iteration (called by ONE thread after every block of the optimiser kernel has finished): clear the next slot of
// the statistics ring, advance the step count and the index-pool cursor.
__device__ __forceinline__ void iter_advance(const IterRef& it, float* stats_base) {
    const int step = it.state[0];
    if (stats_base) { float* nx = stats_base + 4 * ((step + 1) % it.ring); nx[0] = nx[1] = nx[2] = nx[3] = 0.0f; }
    it.state[1] = it.state[1] + 1;
    __threadfence();
    it.state[0] = step + 1;
}

// torch::optim::Adam::step (libtorch defaults, no amsgrad / weight decay) + zero_grad, one launch over the concatenation
// of all segments.  Each thread owns ADAM_VEC float4 slots and issues ALL its loads (g, m, v, p of every slot) before any
// arithmetic, so that the 90 MB pass is bandwidth- rather than latency-bound:
//   m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g g;  p -= (lr / bc1) * m / (sqrt(v)/sqrt(bc2) + eps);  g = 0
constexpr int ADAM_VEC = 2;
// +-0 in all four lanes.  A slot whose gradient, m and v are all zero is left alone: m and v stay 0 and the update is
// p + (-step * 0) / (0 / sqrt(bc2) + eps) = p exactly.  Most of a grid is like that (voxels no ray has reached yet), and those are the
// slow operands of the IEEE divide / sqrt sequences (zero numerators take their out-of-line paths).
__device__ __forceinline__ bool all_zero4(const float4& a) {
    return ((__float_as_uint(a.x) | __float_as_uint(a.y) | __float_as_uint(a.z) | __float_as_uint(a.w)) << 1) == 0u;
}
__device__ __forceinline__ bool all_pos_zero4(const float4& a) {
    return (__float_as_uint(a.x) | __float_as_uint(a.y) | __float_as_uint(a.z) | __float_as_uint(a.w)) == 0u;
}
__global__ void __launch_bounds__(256) k_adam(AdamParams P) {
    const int total = P.cum4[P.n_seg];
    const int t0 = blockIdx.x * (blockDim.x * ADAM_VEC) + threadIdx.x;
    // Per-block table of the segments (first float4, mask, active flag, -lr / bc1): thread k < ADAM_MAX_SEG prepares segment k, so that the
    // double-precision divide, the bias-correction lookup and the statistics-slot modulo run once per block, not once per thread.
    __shared__ int s_begin4[ADAM_MAX_SEG], s_begin[ADAM_MAX_SEG], s_cum[ADAM_MAX_SEG], s_active[ADAM_MAX_SEG];
    __shared__ const uint8_t* s_mask[ADAM_MAX_SEG];
    __shared__ float s_nstep[ADAM_MAX_SEG], s_bc2s;
    __shared__ float* s_loss_dst;
    if (threadIdx.x < ADAM_MAX_SEG) {
        AdamSegment sg = P.seg[0]; int cum = P.cum4[0];
#pragma unroll
        for (int k = 1; k < ADAM_MAX_SEG; ++k)      // compile-time indices only: a runtime index into the by-value parameter struct would go through local memory
            if ((int)threadIdx.x == k) { sg = P.seg[k]; cum = P.cum4[k]; }
        double bc1; float bc2s;
        adam_bias(P, bc1, bc2s);
        s_begin4[threadIdx.x] = sg.begin / 4; s_begin[threadIdx.x] = sg.begin; s_cum[threadIdx.x] = cum; s_active[threadIdx.x] = sg.active; s_mask[threadIdx.x] = sg.mask;
        s_nstep[threadIdx.x] = -(float)((double)sg.lr / bc1);
        if (threadIdx.x == 0) {
            s_bc2s = bc2s;
            s_loss_dst = P.loss_dst ? (P.it.state ? P.loss_dst + 4 * iter_slot(P.it) + 3 : P.loss_dst) : nullptr;
        }
    }
    __syncthreads();
    const float bc2s = s_bc2s;
    float* loss_dst = s_loss_dst;
    int idx[ADAM_VEC]; bool live[ADAM_VEC], upd[ADAM_VEC]; float nstep[ADAM_VEC];
    float4 g[ADAM_VEC], m[ADAM_VEC], v[ADAM_VEC], p[ADAM_VEC];
#pragma unroll
    for (int u = 0; u < ADAM_VEC; ++u) {
        const int tid = t0 + u * blockDim.x;
        live[u] = tid < total;
        int si = 0;
#pragma unroll
        for (int k = 1; k < ADAM_MAX_SEG; ++k) si += (k < P.n_seg && tid >= P.cum4[k]) ? 1 : 0;     // cum4 ascends: the number of segment starts <= tid
        idx[u] = live[u] ? s_begin4[si] + (tid - s_cum[si]) : 0;
        nstep[u] = s_nstep[si];
        const uint8_t* mask = s_mask[si];
        upd[u] = live[u] && s_active[si] && !(mask && !mask[(idx[u] * 4 - s_begin[si]) / CDIM]);
    }
#pragma unroll
    for (int u = 0; u < ADAM_VEC; ++u) {
        if (live[u]) g[u] = reinterpret_cast<const float4*>(P.grad)[idx[u]];
        if (upd[u]) { m[u] = reinterpret_cast<const float4*>(P.m)[idx[u]]; v[u] = reinterpret_cast<const float4*>(P.v)[idx[u]]; p[u] = reinterpret_cast<const float4*>(P.param)[idx[u]]; }
    }
#pragma unroll
    for (int u = 0; u < ADAM_VEC; ++u) {
        if (!live[u]) continue;
        if (loss_dst && idx[u] == P.loss_idx4) *loss_dst = g[u].x;
        if (!all_pos_zero4(g[u])) reinterpret_cast<float4*>(P.grad)[idx[u]] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (!upd[u]) continue;
        if (all_zero4(g[u]) && all_zero4(m[u]) && all_zero4(v[u])) continue;
        float* gg = reinterpret_cast<float*>(&g[u]); float* mm = reinterpret_cast<float*>(&m[u]);
        float* vv = reinterpret_cast<float*>(&v[u]); float* pp = reinterpret_cast<float*>(&p[u]);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float gk = gg[k] * P.grad_scale;
            mm[k] = __fadd_rn(__fmul_rn(mm[k], P.beta1), __fmul_rn(P.om_beta1, gk));                   // mul_(b1).add_(g, 1-b1)
            vv[k] = __fadd_rn(__fmul_rn(vv[k], P.beta2), __fmul_rn(__fmul_rn(P.om_beta2, gk), gk));   // mul_(b2).addcmul_(g, g, 1-b2)
            const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(vv[k]), bc2s), P.eps);
            pp[k] = __fadd_rn(pp[k], __fdiv_rn(__fmul_rn(nstep[u], mm[k]), denom));                   // addcdiv_(m, denom, -step)
        }
        reinterpret_cast<float4*>(P.m)[idx[u]] = m[u]; reinterpret_cast<float4*>(P.v)[idx[u]] = v[u]; reinterpret_cast<float4*>(P.param)[idx[u]] = p[u];
    }
    if (P.it.state) {   // the last block to finish closes the iteration (every block has read state[0] before it arrives here)
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            if (atomicAdd(P.it.state + 2, 1) == (int)gridDim.x - 1) { P.it.state[2] = 0; iter_advance(P.it, P.loss_dst); }
        }
    }
}

// Renderer::eval_points epilogue on the device (Renderer.cpp:26-36 + the stage assembly of NICE.cpp:16-51): raw = (r, g, b, occ)
// with the stage's occupancy sum, and occupancy 100 for points that are not strictly inside the bound.
__global__ void k_assemble_raw(const float* __restrict__ pts, const float* __restrict__ raw_rgb, const float* __restrict__ occ0,
                               const float* __restrict__ occ1, const float* __restrict__ occ2, int stage, Bound bnd, int n,
                               float* __restrict__ raw4, float* __restrict__ occ_only) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float px = pts[3 * (size_t)i], py = pts[3 * (size_t)i + 1], pz = pts[3 * (size_t)i + 2];
    float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
    if (stage == 3) { const float4 c = *reinterpret_cast<const float4*>(raw_rgb + 4 * (size_t)i); r.x = c.x; r.y = c.y; r.z = c.z; }
    const float occ = stage == 0 ? occ0[i] : stage == 1 ? occ1[i] : __fadd_rn(occ2[i], occ1[i]);
    const bool in = px < bnd.hi[0] && px > bnd.lo[0] && py < bnd.hi[1] && py > bnd.lo[1] && pz < bnd.hi[2] && pz > bnd.lo[2];
    r.w = in ? occ : 100.0f;
    if (raw4) *reinterpret_cast<float4*>(raw4 + 4 * (size_t)i) = r;
    if (occ_only) occ_only[i] = r.w;
}

// Points of a regular lattice, generated on the device (mesh extraction queries, nice_slam.yaml meshing.resolution): point index
// q = (j * nx + i) * nz + k  <->  (x_i, y_j, z_k), the ravel order of numpy.meshgrid(x, y, z) that upstream's mesher uses;
// axis values are torch/numpy linspace(lo, hi, n) in fp32 (symmetric formula).
struct LatticeParams { int nx, ny, nz; float lo[3], hi[3]; };
__device__ __forceinline__ float lattice_axis(float lo, float hi, int n, int i) {
    if (n <= 1) return lo;
    const float step = __fdiv_rn(__fsub_rn(hi, lo), (float)(n - 1));
    return i < n / 2 ? __fadd_rn(lo, __fmul_rn(step, (float)i)) : __fsub_rn(hi, __fmul_rn(step, (float)(n - 1 - i)));
}
__global__ void k_lattice_points(LatticeParams L, long long first, int n, float* __restrict__ pts) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const long long q = first + t;
    const int k = (int)(q % L.nz); const long long r = q / L.nz;
    const int i = (int)(r % L.nx), j = (int)(r / L.nx);
    pts[3 * (size_t)t] = lattice_axis(L.lo[0], L.hi[0], L.nx, i);
    pts[3 * (size_t)t + 1] = lattice_axis(L.lo[1], L.hi[1], L.ny, j);
    pts[3 * (size_t)t + 2] = lattice_axis(L.lo[2], L.hi[2], L.nz, k);
}

__global__ void k_fill(float* p, float v, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

}  // namespace nsb
