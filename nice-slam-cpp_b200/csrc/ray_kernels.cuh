// Per-ray kernels: pixel -> ray generation + inside filter, depth-guided sample placement, alpha compositing
// (forward and backward), losses, lower-median select, pose-gradient reduction.
// Replaces raySampler/get_samples (utils.h:13-55,141-146), the filter of Mapper.cpp:416-427 / Tracker.cpp:48-58,
// Renderer.cpp:46-119 (z values), raw2outputs_nerf_color (utils.h:148-172), the losses of Mapper.cpp:435-442 and
// Tracker.cpp:67-82 and the autograd backward of all of them.
#pragma once
#include "common.cuh"

namespace nsb {

constexpr int MAX_OPT_FRAMES = 16;

// ---- utils.h:174-195 -------------------------------------------------------------------------------------
__host__ __device__ inline void quad2rotation(const float* q, float* R) {
    const float qr = q[0], qi = q[1], qj = q[2], qk = q[3];
    const float two_s = 2.0f / (qr * qr + qi * qi + qj * qj + qk * qk);
    R[0] = 1 - two_s * (qj * qj + qk * qk); R[1] = two_s * (qi * qj - qk * qr); R[2] = two_s * (qi * qk + qj * qr);
    R[3] = two_s * (qi * qj + qk * qr); R[4] = 1 - two_s * (qi * qi + qk * qk); R[5] = two_s * (qj * qk - qi * qr);
    R[6] = two_s * (qi * qk - qj * qr); R[7] = two_s * (qj * qk + qi * qr); R[8] = 1 - two_s * (qi * qi + qj * qj);
}

// d L / d q given G = d L / d R (row-major 3x3): R = I + two_s * A(q), two_s = 2 / |q|^2.
__host__ __device__ inline void quad2rotation_vjp(const float* q, const float* G, float* gq) {
    const float qr = q[0], qi = q[1], qj = q[2], qk = q[3];
    const float n = qr * qr + qi * qi + qj * qj + qk * qk, two_s = 2.0f / n;
    const float A[9] = {-(qj * qj + qk * qk), qi * qj - qk * qr, qi * qk + qj * qr,
                        qi * qj + qk * qr, -(qi * qi + qk * qk), qj * qk - qi * qr,
                        qi * qk - qj * qr, qj * qk + qi * qr, -(qi * qi + qj * qj)};
    float GA = 0.0f;
    for (int i = 0; i < 9; ++i) GA += G[i] * A[i];
    // dA/dq_m contracted with G
    const float dqr = G[1] * (-qk) + G[2] * qj + G[3] * qk + G[5] * (-qi) + G[6] * (-qj) + G[7] * qi;
    const float dqi = G[1] * qj + G[2] * qk + G[3] * qj + G[4] * (-2 * qi) + G[5] * (-qr) + G[6] * qk + G[7] * qr + G[8] * (-2 * qi);
    const float dqj = G[0] * (-2 * qj) + G[1] * qi + G[2] * qr + G[3] * qi + G[5] * qk + G[6] * (-qr) + G[7] * qk + G[8] * (-2 * qj);
    const float dqk = G[0] * (-2 * qk) + G[1] * (-qr) + G[2] * qi + G[3] * qr + G[4] * (-2 * qk) + G[5] * qj + G[6] * qi + G[7] * qj;
    const float k = two_s * two_s * GA;   // d two_s / d q_m = -two_s^2 q_m
    gq[0] = two_s * dqr - k * qr; gq[1] = two_s * dqi - k * qi; gq[2] = two_s * dqj - k * qj; gq[3] = two_s * dqk - k * qk;
}

// Multi-GPU ray order.  The reference's batch is frame-major: ray q of frame f is element f * pix + q (Mapper.cpp:404-414).
// With `world` ranks each rank renders a contiguous block of `per` rays; to give every rank the same mix of frames (same
// inside fraction, same scatter footprint) the blocks are cut rank-major: element i of the rendered order is
//   rank r = i / per, j = i % per, frame f = j / ppr, k = j % ppr   ->   reference element  f * pix + r * ppr + k,   ppr = pix / world.
// The loss and every gradient are sums over rays, so the order only changes the association of fp32 additions.
struct RayOrder {
    int world, per, pix, ppr;    // world <= 1 (incl. a zeroed struct): identity, frame = i / pix_per_frame
    __host__ __device__ __forceinline__ bool identity() const { return world <= 1; }
    __host__ __device__ __forceinline__ int frame(int i, int pix_per_frame) const { return identity() ? i / pix_per_frame : (i % per) / ppr; }
    __host__ __device__ __forceinline__ int source(int i) const {
        if (identity()) return i;
        const int r = i / per, j = i % per;
        return (j / ppr) * pix + r * ppr + (j % ppr);
    }
};

struct SampleParams {
    const float* depth;      // [max_frames][H*W]
    const float* color;      // [max_frames][H*W*3]
    const float* poses;      // [max_frames][12]  row-major [R|t]
    const float* cams;       // [n_frames][8] 7-vector poses being optimised (tracking: one; bundle adjustment: one per frame)
    uint32_t cam_mask;       // bit f set: frame f takes its pose from cams + 8 f (R via quad2rotation) instead of poses[slot]
    const int64_t* idx;      // [n] flat index into the crop
    const int64_t* pool;     // optional resident index pool [pool_iters][n]: row (it.state[1] % pool_iters) replaces idx
    int pool_iters;
    IterRef it;              // mapping loop: selects the statistics slot (and the pool row); {nullptr, 1} elsewhere
    int slots[MAX_OPT_FRAMES];
    int n_frames, pix_per_frame;
    RayOrder order;          // which reference batch element ray i is (identity on one GPU)
    int H, W, H0, W0, Wc;
    float fx, fy, cx, cy;
    int raydir;
    Bound bnd;
    int n;
    float* rays_o; float* rays_d; float* gt_depth; float* gt_color; uint8_t* valid;
    float* stats;            // base of the statistics ring; slot: [0] max gt_depth over valid rays (as int bits), [1] count of valid rays (int)
    int apply_filter;        // 0: valid = 1 for every ray (render API)
};

__device__ __forceinline__ float aabb_exit(const Bound& b, const float* o, const float* d) {
    float t = INFINITY;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float t0 = __fdiv_rn(__fsub_rn(b.lo[a], o[a]), d[a]);
        const float t1 = __fdiv_rn(__fsub_rn(b.hi[a], o[a]), d[a]);
        t = fminf(t, fmaxf(t0, t1));
    }
    return t;
}

__device__ __forceinline__ const int64_t* pool_row(const int64_t* idx, const int64_t* pool, int pool_iters, const IterRef& it, int n) {
    return pool ? pool + (size_t)((it.state ? it.state[1] : 0) % pool_iters) * n : idx;
}

__global__ void k_sample(SampleParams P) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    float gd = 0.0f; bool ok = false;
    float* stats = P.stats + 4 * iter_slot(P.it);
    if (i < P.n) {
        const int f = min(P.order.frame(i, P.pix_per_frame), P.n_frames - 1);
        const int slot = P.slots[f];
        const int64_t id = pool_row(P.idx, P.pool, P.pool_iters, P.it, P.n)[P.order.source(i)];
        const int x = P.W0 + (int)(id % P.Wc), y = P.H0 + (int)(id / P.Wc);
        const size_t pix = (size_t)slot * P.H * P.W + (size_t)y * P.W + x;
        gd = P.depth[pix];
        const float* c = P.color + pix * 3;
        P.gt_color[3 * i + 0] = c[0]; P.gt_color[3 * i + 1] = c[1]; P.gt_color[3 * i + 2] = c[2];
        P.gt_depth[i] = gd;
        float R[9], tr[3];
        if ((P.cam_mask >> f) & 1u) {
            const float* q = P.cams + 8 * f;
            quad2rotation(q, R);
            tr[0] = q[4]; tr[1] = q[5]; tr[2] = q[6];
        } else {
            const float* m = P.poses + slot * 12;
            for (int r = 0; r < 3; ++r) { R[3 * r] = m[4 * r]; R[3 * r + 1] = m[4 * r + 1]; R[3 * r + 2] = m[4 * r + 2]; tr[r] = m[4 * r + 3]; }
        }
        // utils.h:44-47 (reference: j_t uses i and is not negated) or upstream pinhole
        const float xf = (float)x, yf = (float)y;
        const float d0 = __fdiv_rn(__fsub_rn(xf, P.cx), P.fx);
        const float d1 = P.raydir == 0 ? __fdiv_rn(__fsub_rn(xf, P.cy), P.fy) : -__fdiv_rn(__fsub_rn(yf, P.cy), P.fy);
        const float d2 = -1.0f;
        float o[3], d[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) {   // utils.h:51: sum(dirs * c2w[:3,:3], -1)
            d[r] = __fadd_rn(__fadd_rn(__fmul_rn(d0, R[3 * r]), __fmul_rn(d1, R[3 * r + 1])), __fmul_rn(d2, R[3 * r + 2]));
            o[r] = tr[r];
            P.rays_d[3 * i + r] = d[r]; P.rays_o[3 * i + r] = o[r];
        }
        ok = P.apply_filter ? (aabb_exit(P.bnd, o, d) >= gd) : true;   // Mapper.cpp:420-423
        P.valid[i] = ok ? 1 : 0;
    }
    // batch-global scalars of Renderer.cpp:76,93 are taken over the rays that pass the filter
    const unsigned bal = __ballot_sync(0xffffffffu, ok);
    float m = ok ? gd : 0.0f;
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, s));
    if ((threadIdx.x & 31) == 0 && bal) {
        atomicMax(reinterpret_cast<int*>(stats), __float_as_int(m));
        atomicAdd(reinterpret_cast<int*>(stats) + 1, __popc(bal));
    }
}

// max gt_depth for the direct render API (rays supplied by the caller)
__global__ void k_depth_max(const float* __restrict__ gt_depth, int n, float* stats) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    float m = i < n ? fmaxf(gt_depth[i], 0.0f) : 0.0f;
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, s));
    if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<int*>(stats), __float_as_int(m));
}

// utils.h:153 as written: torch::norm(x, -1) = (sum |x|^-1)^-1 over the whole batch.  stats[2] accumulates sum 1/|d_ij|.
__global__ void k_dirnorm_ref(const float* __restrict__ rays_d, const uint8_t* __restrict__ valid, int n, float* stats, IterRef it) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    stats += 4 * iter_slot(it);
    float s = 0.0f;
    if (i < n && (!valid || valid[i])) s = 1.0f / fabsf(rays_d[3 * i]) + 1.0f / fabsf(rays_d[3 * i + 1]) + 1.0f / fabsf(rays_d[3 * i + 2]);
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) atomicAdd(stats + 2, s);
}

struct ZParams {
    const float* rays_o; const float* rays_d; const float* gt_depth;   // gt_depth null: no-depth path
    const uint8_t* valid;
    const float* stats;      // statistics ring base; slot[0] = max gt_depth
    IterRef it;
    unsigned long long* zero_ctr;   // tile counters of the forward decoder launch that follows (cleared here), or nullptr
    const float* t_samples;  // [32]
    const float* t_surface;  // [16]
    Bound bnd;
    int n, n_samples, n_surface;
    float* z;                // [n][S]
    int* ray_list;           // or nullptr: compacted list of the rays that pass the inside filter (any order), for the tcgen05 forward
    int* ray_count;          // its length (zeroed by the forward kernel's last CTA)
};

// Renderer.cpp:46-119: one warp per ray.  32 stratified + 16 near-surface values, then an exact rank sort.
__global__ void k_zvals(ZParams P) {
    const int ray = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, l = threadIdx.x & 31;
    if (blockIdx.x == 0 && threadIdx.x == 0) zero_tile_counters(P.zero_ctr);
    const bool live = ray < P.n && !(P.valid && !P.valid[ray]);
    if (P.ray_list) {   // one global atomic per block: the block's surviving rays take consecutive slots of the list
        __shared__ int s_cnt, s_base;
        if (threadIdx.x == 0) s_cnt = 0;
        __syncthreads();
        int pos = 0;
        if (live && l == 0) pos = atomicAdd(&s_cnt, 1);
        __syncthreads();
        if (threadIdx.x == 0 && s_cnt > 0) s_base = atomicAdd(P.ray_count, s_cnt);
        __syncthreads();
        if (live && l == 0) P.ray_list[s_base + pos] = ray;
    }
    if (!live) return;
    const float o[3] = {P.rays_o[3 * ray], P.rays_o[3 * ray + 1], P.rays_o[3 * ray + 2]};
    const float d[3] = {P.rays_d[3 * ray], P.rays_d[3 * ray + 1], P.rays_d[3 * ray + 2]};
    // :69-73, aabb_exit with one IEEE divide per lane instead of six per lane: lane a < 6 forms (bound - o) / d of axis a % 3 (lo: a < 3, hi: a >= 3)
    float far_bb;
    {
        const int ax = l % 3;
        const float bv = (l % 6) < 3 ? (ax == 0 ? P.bnd.lo[0] : ax == 1 ? P.bnd.lo[1] : P.bnd.lo[2]) : (ax == 0 ? P.bnd.hi[0] : ax == 1 ? P.bnd.hi[1] : P.bnd.hi[2]);
        const float oa = ax == 0 ? o[0] : ax == 1 ? o[1] : o[2], da = ax == 0 ? d[0] : ax == 1 ? d[1] : d[2];
        const float tq = __fdiv_rn(__fsub_rn(bv, oa), da);
        float te = INFINITY;
#pragma unroll
        for (int a = 0; a < 3; ++a) te = fminf(te, fmaxf(__shfl_sync(0xffffffffu, tq, a), __shfl_sync(0xffffffffu, tq, a + 3)));
        far_bb = __fadd_rn(te, 0.01f);
    }
    const float t = P.t_samples[l], u = __fsub_rn(1.0f, t);
    if (!P.gt_depth) {                                                               // :54-58,78,106
        P.z[ray * P.n_samples + l] = __fadd_rn(__fmul_rn(0.01f, u), __fmul_rn(far_bb, t));
        return;
    }
    const float gd = P.gt_depth[ray], gmax = P.stats[4 * iter_slot(P.it)];
    const float near = __fmul_rn(gd, 0.01f);                                         // :63
    const float far = fminf(fmaxf(far_bb, 0.0f), __fmul_rn(gmax, 1.2f));            // :76
    const float v0 = __fadd_rn(__fmul_rn(near, u), __fmul_rn(far, t));              // :106
    float v1 = 0.0f;
    const int S = P.n_samples + P.n_surface;
    if (l < P.n_surface) {
        const float ts = P.t_surface[l], us = __fsub_rn(1.0f, ts);
        if (gd > 0.0f) v1 = __fadd_rn(__fmul_rn(__fmul_rn(0.95f, gd), us), __fmul_rn(__fmul_rn(1.05f, gd), ts));   // :88
        else v1 = __fadd_rn(__fmul_rn(0.001f, us), __fmul_rn(gmax, ts));             // :94
    }
    if (P.n_surface == 0) { P.z[ray * S + l] = v0; return; }
    // :119 sort == rank placement (stable: ties keep the concatenation order [stratified | surface]).  Both runs ascend by construction
    // (t ascending, near <= far), so a value's rank is its own index plus the number of values of the OTHER run that go before it -- a
    // binary search over the other run's lanes.  If a run does not ascend (or holds a NaN) the plain O(S) count below decides.
    int r0, r1;
    {
        const float p0 = __shfl_up_sync(0xffffffffu, v0, 1), p1 = __shfl_up_sync(0xffffffffu, v1, 1);
        const bool asc = (l == 0 || p0 <= v0) && (l == 0 || l >= P.n_surface || p1 <= v1) && P.n_samples == 32;
        if (__all_sync(0xffffffffu, asc)) {
            int c0 = 0, c1 = 0;          // c0 = #{j < n_surface : v1[j] < v0},  c1 = #{k < 32 : v0[k] <= v1}
#pragma unroll
            for (int step = 32; step > 0; step >>= 1) {
                const float e1 = __shfl_sync(0xffffffffu, v1, (c0 + step - 1) & 31);
                const float e0 = __shfl_sync(0xffffffffu, v0, (c1 + step - 1) & 31);
                if (c0 + step <= P.n_surface && e1 < v0) c0 += step;
                if (c1 + step <= 32 && e0 <= v1) c1 += step;
            }
            r0 = l + c0; r1 = l + c1;
        } else {
            r0 = 0; r1 = 0;
            for (int k = 0; k < S; ++k) {
                const float uk = k < 32 ? __shfl_sync(0xffffffffu, v0, k) : __shfl_sync(0xffffffffu, v1, k - 32);
                r0 += (uk < v0) || (uk == v0 && k < l);
                r1 += (uk < v1) || (uk == v1 && k < 32 + l);
            }
        }
    }
    P.z[ray * S + r0] = v0;
    if (l < P.n_surface) P.z[ray * S + r1] = v1;
}

// ---- compositing -------------------------------------------------------------------------------------
struct CompositeParams {
    const float* rays_o; const float* rays_d; const float* z; const uint8_t* valid;
    const float* raw_rgb;     // [P][4]
    const float* occ[3];      // coarse, middle, fine
    const float* stats;       // statistics ring base; slot[2] = sum 1/|d| for the reference dist norm
    IterRef it;
    unsigned long long* zero_ctr;   // tile counters of the backward decoder launch that follows (cleared here), or nullptr
    Bound bnd;
    int n, S, stage, occupancy, dist_norm;
    // forward outputs
    float* rgb; float* depth; float* var; float* weights;
    // backward
    const float* g_rgb; const float* g_depth; const float* g_var;   // per-ray cotangents
    float* g_raw;             // [P][4]
    float* d_rays;            // [n][6] or null: receives d L / d rays_d through dists * |rays_d| (utils.h:153)
    // tracking (Tracker.cpp:69): the forward also compacts |gt_depth - depth| of the surviving rays for the median
    const float* trk_gt_depth; float* trk_absdiff; int* trk_count;
};

struct RaySamples {   // lane l owns samples l and l+32
    float z[2], sig[2], col[2][3], delta[2], alpha[2], T[2], w[2];
    bool in[2], has[2];
};

__device__ __forceinline__ float warp_scan_mul(float v, int l) {   // inclusive prefix product
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const float n = __shfl_up_sync(0xffffffffu, v, o);
        if (l >= o) v *= n;
    }
    return v;
}
__device__ __forceinline__ float warp_scan_add_rev(float v, int l) {   // inclusive suffix sum
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const float n = __shfl_down_sync(0xffffffffu, v, o);
        if (l + o < 32) v += n;
    }
    return v;
}

// Loads the ray's samples and runs the forward composite (utils.h:150-171).  Returns rgb/depth/var in all lanes.
__device__ __forceinline__ void composite_forward(const CompositeParams& P, int ray, int l, RaySamples& s,
                                                  float (&rgb)[3], float& depth, float& var) {
    const float o[3] = {P.rays_o[3 * ray], P.rays_o[3 * ray + 1], P.rays_o[3 * ray + 2]};
    const float d[3] = {P.rays_d[3 * ray], P.rays_d[3 * ray + 1], P.rays_d[3 * ray + 2]};
    float norm;
    if (P.dist_norm == 0) norm = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    else norm = 1.0f / P.stats[4 * iter_slot(P.it) + 2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int k = l + 32 * h;
        s.has[h] = k < P.S;
        s.z[h] = 0.0f; s.sig[h] = 0.0f; s.in[h] = false; s.delta[h] = 0.0f;
        s.col[h][0] = s.col[h][1] = s.col[h][2] = 0.0f;
        if (s.has[h]) {
            const int q = ray * P.S + k;
            const float z = P.z[q];
            s.z[h] = z;
            const float zn = k + 1 < P.S ? P.z[q + 1] : 0.0f;
            s.delta[h] = __fmul_rn(k + 1 < P.S ? __fsub_rn(zn, z) : 1e10f, norm);
            float sig;
            if (P.stage == 0) sig = P.occ[0][q];
            else if (P.stage == 1) sig = P.occ[1][q];
            else sig = __fadd_rn(P.occ[2][q], P.occ[1][q]);                        // NICE.cpp:40,49
            if (P.stage == 3) { const float4 c = *reinterpret_cast<const float4*>(P.raw_rgb + 4 * (size_t)q); s.col[h][0] = c.x; s.col[h][1] = c.y; s.col[h][2] = c.z; }
            bool in = true;                                                         // Renderer.cpp:26-29
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const float p = __fadd_rn(o[a], __fmul_rn(d[a], z));               // Renderer.cpp:121
                in = in && (p < P.bnd.hi[a]) && (p > P.bnd.lo[a]);
            }
            s.in[h] = in;
            s.sig[h] = in ? sig : 100.0f;                                           // Renderer.cpp:36
        }
        if (P.occupancy) s.alpha[h] = s.has[h] ? 1.0f / (1.0f + expf(-10.0f * s.sig[h])) : 0.0f;
        else s.alpha[h] = s.has[h] ? 1.0f - expf(-(fmaxf(s.sig[h], 0.0f) * s.delta[h])) : 0.0f;
    }
    // transmittance: exclusive prefix product of (1 - alpha + 1e-10)  (utils.h:159-164)
    const float q0 = __fadd_rn(__fsub_rn(1.0f, s.alpha[0]), 1e-10f);
    const float q1 = s.has[1] ? __fadd_rn(__fsub_rn(1.0f, s.alpha[1]), 1e-10f) : 1.0f;
    const float inc0 = warp_scan_mul(q0, l);
    float ex0 = __shfl_up_sync(0xffffffffu, inc0, 1); if (l == 0) ex0 = 1.0f;
    const float tot0 = __shfl_sync(0xffffffffu, inc0, 31);
    const float inc1 = warp_scan_mul(q1, l);
    float ex1 = __shfl_up_sync(0xffffffffu, inc1, 1); if (l == 0) ex1 = 1.0f;
    s.T[0] = ex0; s.T[1] = tot0 * ex1;
    s.w[0] = s.alpha[0] * s.T[0];
    s.w[1] = s.has[1] ? s.alpha[1] * s.T[1] : 0.0f;
    rgb[0] = warp_sum(s.w[0] * s.col[0][0] + s.w[1] * s.col[1][0]);
    rgb[1] = warp_sum(s.w[0] * s.col[0][1] + s.w[1] * s.col[1][1]);
    rgb[2] = warp_sum(s.w[0] * s.col[0][2] + s.w[1] * s.col[1][2]);
    depth = warp_sum(s.w[0] * s.z[0] + s.w[1] * s.z[1]);
    const float t0 = s.z[0] - depth, t1 = s.z[1] - depth;
    var = warp_sum(s.w[0] * t0 * t0 + s.w[1] * t1 * t1);
}

__global__ void k_composite_fwd(CompositeParams P) {
    const int ray = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, l = threadIdx.x & 31;
    const bool live = ray < P.n && !(P.valid && !P.valid[ray]);
    float absd = 0.0f;
    if (ray < P.n && !live) {
        if (l == 0) { P.rgb[3 * ray] = P.rgb[3 * ray + 1] = P.rgb[3 * ray + 2] = 0.0f; P.depth[ray] = 0.0f; P.var[ray] = 0.0f; }
        if (P.weights) for (int k = l; k < P.S; k += 32) P.weights[ray * P.S + k] = 0.0f;
    } else if (live) {
        RaySamples s; float rgb[3], depth, var;
        composite_forward(P, ray, l, s, rgb, depth, var);
        if (l == 0) { P.rgb[3 * ray] = rgb[0]; P.rgb[3 * ray + 1] = rgb[1]; P.rgb[3 * ray + 2] = rgb[2]; P.depth[ray] = depth; P.var[ray] = var; }
        if (P.trk_absdiff && l == 0) absd = fabsf(P.trk_gt_depth[ray] - depth);
        if (P.weights) {
            P.weights[ray * P.S + l] = s.w[0];
            if (s.has[1]) P.weights[ray * P.S + 32 + l] = s.w[1];
        }
    }
    if (P.trk_absdiff) {   // tracking: |gt_depth - depth| of the surviving rays, compacted for the median (any order); one global atomic per block
        __shared__ int s_cnt, s_base;
        if (threadIdx.x == 0) s_cnt = 0;
        __syncthreads();
        int pos = 0;
        if (live && l == 0) pos = atomicAdd(&s_cnt, 1);
        __syncthreads();
        if (threadIdx.x == 0 && s_cnt > 0) s_base = atomicAdd(P.trk_count, s_cnt);
        __syncthreads();
        if (live && l == 0) P.trk_absdiff[s_base + pos] = absd;
    }
}

// Backward of the composite for the ray's cotangents (gC, gD0, gV) -> g_raw (r,g,b,occ) per sample.
__device__ __forceinline__ void composite_backward(const CompositeParams& P, int ray, int l, const RaySamples& s, float depth,
                                                   const float (&gC)[3], float gD0, float gV) {
    // var = sum w (z - D)^2  =>  dvar/dD = -2 sum w (z - D)
    const float sw = warp_sum(s.w[0] * (s.z[0] - depth) + s.w[1] * (s.z[1] - depth));
    const float gD = gD0 - 2.0f * gV * sw;
    float gw[2], gwsum[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const float dz = s.z[h] - depth;
        gw[h] = s.has[h] ? gC[0] * s.col[h][0] + gC[1] * s.col[h][1] + gC[2] * s.col[h][2] + gD * s.z[h] + gV * dz * dz : 0.0f;
    }
    // suffix sums of gw_k * w_k over k > j  (cumprod backward: d T_k / d q_j = T_k / q_j)
    const float a0 = gw[0] * s.w[0], a1 = gw[1] * s.w[1];
    const float suf1 = warp_scan_add_rev(a1, l), tot1 = __shfl_sync(0xffffffffu, suf1, 0);
    const float suf0 = warp_scan_add_rev(a0, l);
    gwsum[0] = (suf0 - a0) + tot1;
    gwsum[1] = suf1 - a1;
    float g_norm = 0.0f;   // d L / d |rays_d| through delta = dz * |rays_d|
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        if (!s.has[h]) continue;
        const int q = ray * P.S + l + 32 * h;
        const float qq = __fadd_rn(__fsub_rn(1.0f, s.alpha[h]), 1e-10f);
        const float g_alpha = gw[h] * s.T[h] - gwsum[h] / qq;
        if (!P.occupancy && s.sig[h] > 0.0f) {
            const float e = expf(-(s.sig[h] * s.delta[h]));
            if (e > 0.0f) g_norm += g_alpha * s.sig[h] * e * s.delta[h];   // times 1/|d| below (delta = dz |d|)
        }
        float g_sig;
        if (P.occupancy) g_sig = g_alpha * 10.0f * s.alpha[h] * (1.0f - s.alpha[h]);
        else g_sig = s.sig[h] > 0.0f ? g_alpha * s.delta[h] * expf(-(s.sig[h] * s.delta[h])) : 0.0f;
        if (!s.in[h]) g_sig = 0.0f;   // occupancy overwritten with 100 (Renderer.cpp:36): no gradient to the decoders
        float4 o;
        o.x = s.w[h] * gC[0]; o.y = s.w[h] * gC[1]; o.z = s.w[h] * gC[2]; o.w = g_sig;
        if (P.stage != 3) { o.x = o.y = o.z = 0.0f; }
        *reinterpret_cast<float4*>(P.g_raw + 4 * (size_t)q) = o;
    }
    if (P.d_rays && P.dist_norm == 0) {
        g_norm = warp_sum(g_norm);
        if (l < 3) {
            const float dx = P.rays_d[3 * ray], dy = P.rays_d[3 * ray + 1], dz = P.rays_d[3 * ray + 2];
            const float n2 = dx * dx + dy * dy + dz * dz;   // d|d|/d d = d/|d|, and delta/|d| = dz
            atomicAdd(P.d_rays + 6 * (size_t)ray + 3 + l, g_norm * P.rays_d[3 * ray + l] / n2);
        }
    }
}

__global__ void k_composite_bwd(CompositeParams P) {
    const int ray = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, l = threadIdx.x & 31;
    if (blockIdx.x == 0 && threadIdx.x == 0) zero_tile_counters(P.zero_ctr);
    if (ray >= P.n) return;
    if (P.valid && !P.valid[ray]) return;   // tiles of dropped rays are skipped by the decoder kernels too
    RaySamples s; float rgb[3], depth, var;
    composite_forward(P, ray, l, s, rgb, depth, var);
    const float gC[3] = {P.g_rgb[3 * ray], P.g_rgb[3 * ray + 1], P.g_rgb[3 * ray + 2]};
    composite_backward(P, ray, l, s, depth, gC, P.g_depth[ray], P.g_var[ray]);
}

// Mapping iteration: composite + loss of Mapper.cpp:435-442 + composite backward in ONE pass over the ray (the loss
// cotangents are local to the ray).  Writes rgb / depth / var, accumulates the loss, leaves g_raw for the decoder backward.
__global__ void k_composite_map(CompositeParams P, const float* __restrict__ gt_depth, const float* __restrict__ gt_color,
                                int use_color, float w_color, float* loss_out) {
    const int ray = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, l = threadIdx.x & 31;
    if (blockIdx.x == 0 && threadIdx.x == 0) zero_tile_counters(P.zero_ctr);
    __shared__ float s_loss[32];              // per-warp loss terms: one global atomic per block instead of one per ray
    float loss = 0.0f;
    if (ray < P.n && P.valid && !P.valid[ray]) {
        if (l == 0) { P.rgb[3 * ray] = P.rgb[3 * ray + 1] = P.rgb[3 * ray + 2] = 0.0f; P.depth[ray] = 0.0f; P.var[ray] = 0.0f; }
    } else if (ray < P.n) {
        RaySamples s; float rgb[3], depth, var;
        composite_forward(P, ray, l, s, rgb, depth, var);
        float gD = 0.0f, gC[3] = {0.0f, 0.0f, 0.0f};
        const float g = gt_depth[ray];
        if (g > 0.0f) { const float df = g - depth; loss += fabsf(df); gD = df > 0.0f ? -1.0f : (df < 0.0f ? 1.0f : 0.0f); }
        if (use_color) {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float df = gt_color[3 * ray + c] - rgb[c];
                loss += w_color * fabsf(df);
                gC[c] = df > 0.0f ? -w_color : (df < 0.0f ? w_color : 0.0f);
            }
        }
        if (l == 0) { P.rgb[3 * ray] = rgb[0]; P.rgb[3 * ray + 1] = rgb[1]; P.rgb[3 * ray + 2] = rgb[2]; P.depth[ray] = depth; P.var[ray] = var; }
        composite_backward(P, ray, l, s, depth, gC, gD, 0.0f);
    }
    if (l == 0) s_loss[threadIdx.x >> 5] = loss;
    __syncthreads();
    if (threadIdx.x < 32) {
        const float v = warp_sum(threadIdx.x < (blockDim.x >> 5) ? s_loss[threadIdx.x] : 0.0f);
        if (threadIdx.x == 0 && v != 0.0f) atomicAdd(loss_out, v);
    }
}

// Tracking iteration: loss of Tracker.cpp:67-82 (median mask, uncertainty-weighted depth term, colour term) + composite backward
// in one pass over the ray, from the rgb / depth / var the forward left (the cotangents are local to the ray).
__global__ void k_composite_track(CompositeParams P, const float* __restrict__ gt_depth, const float* __restrict__ gt_color,
                                  const float* __restrict__ median, int handle_dynamic, int use_color, float w_color, float* loss_out) {
    const int ray = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, l = threadIdx.x & 31;
    if (blockIdx.x == 0 && threadIdx.x == 0) zero_tile_counters(P.zero_ctr);
    __shared__ float s_loss[32];              // per-warp loss terms: one global atomic per block
    float loss = 0.0f;
    if (ray < P.n && !(P.valid && !P.valid[ray])) {
        RaySamples s; float rgb[3], depth, var;
        composite_forward(P, ray, l, s, rgb, depth, var);
        float gD = 0.0f, gV = 0.0f, gC[3] = {0.0f, 0.0f, 0.0f};
        const float g = gt_depth[ray], df = g - depth;
        bool m = g > 0.0f;
        if (handle_dynamic) m = m && (fabsf(df) < 10.0f * median[0]);
        if (m) {
            const float vv = var + 1e-10f, r = 1.0f / sqrtf(vv);
            loss += fabsf(df) * r;
            gD = -(df > 0.0f ? 1.0f : (df < 0.0f ? -1.0f : 0.0f)) * r;
            gV = -0.5f * fabsf(df) * r / vv;
            if (use_color) {
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const float dc = gt_color[3 * ray + c] - rgb[c];
                    loss += w_color * fabsf(dc);
                    gC[c] = dc > 0.0f ? -w_color : (dc < 0.0f ? w_color : 0.0f);
                }
            }
        }
        composite_backward(P, ray, l, s, depth, gC, gD, gV);
    }
    if (l == 0) s_loss[threadIdx.x >> 5] = loss;
    __syncthreads();
    if (threadIdx.x < 32) {
        const float v = warp_sum(threadIdx.x < (blockDim.x >> 5) ? s_loss[threadIdx.x] : 0.0f);
        if (threadIdx.x == 0 && v != 0.0f) atomicAdd(loss_out, v);
    }
}

// ---- lower median of the tracking residuals (Tracker.cpp:70) ------------------------------------------
// torch.median (lower median) of n non-negative floats by 4-pass radix select on the bit pattern; one block.
__global__ void k_median(const float* __restrict__ v, const int* __restrict__ count, float* out) {
    __shared__ unsigned hist[256];
    __shared__ unsigned prefix_s, want_s;
    const int n = *count;
    if (n <= 0) { if (threadIdx.x == 0) out[0] = 0.0f; return; }
    if (threadIdx.x == 0) { prefix_s = 0; want_s = (unsigned)((n - 1) / 2); }
    __syncthreads();
    for (int pass = 0; pass < 4; ++pass) {
        const int shift = 24 - 8 * pass;
        for (int b = threadIdx.x; b < 256; b += blockDim.x) hist[b] = 0;
        __syncthreads();
        const unsigned prefix = prefix_s;
        const unsigned mask = pass == 0 ? 0u : (0xffffffffu << (shift + 8));
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const unsigned u = __float_as_uint(v[i]);
            if ((u & mask) == prefix) atomicAdd(&hist[(u >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (threadIdx.x < 32) {   // first bin whose cumulative count exceeds `want` (bin 255 at the latest): lane = 8 bins, warp scan over the lanes
            const int lane = threadIdx.x;
            const unsigned want = want_s;
            unsigned h[8], sum = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) { h[j] = hist[8 * lane + j]; sum += h[j]; }
            unsigned inc = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const unsigned nb = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += nb; }
            const unsigned exc = inc - sum;
            const unsigned hit = __ballot_sync(0xffffffffu, exc <= want && want < inc);
            const int sel = hit ? __ffs(hit) - 1 : 31;
            if (lane == sel) {
                unsigned w = want - exc; int j = 0; bool done = false;
#pragma unroll
                for (int k = 0; k < 7; ++k) if (!done) { if (h[k] <= w) { w -= h[k]; j = k + 1; } else done = true; }
                want_s = w; prefix_s = prefix | ((unsigned)(8 * lane + j) << shift);
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = __uint_as_float(prefix_s);
}

// ---- pose gradient: d L / d (q, t) from the per-ray gradients (block reduction, one block) -------------
struct PoseGradParams {
    const float* d_rays;     // [n][6]: d L / d rays_o, d L / d rays_d
    const int64_t* idx; const uint8_t* valid;
    const int64_t* pool; int pool_iters; int n_total; IterRef it;   // resident index pool (see SampleParams), n_total = rays per row
    const float* cams;       // [n_frames][8] current 7-vectors
    uint32_t cam_mask;       // frames whose pose is optimised
    int pix_per_frame, n_frames;
    RayOrder order;
    int lo, hi;              // rays [lo, hi) carry gradient on this rank (the others' partial sums arrive through the all-reduce)
    int H0, W0, Wc, raydir;
    float fx, fy, cx, cy;
    float* g_cams;           // [n_frames][8]
};

// One block per frame: d L / d (q, t) of that frame's pose from the ray gradients of its pixels
// (rays_o = t, rays_d = R(q) dir: utils.h:44-52, 174-210).
__global__ void k_pose_grad(PoseGradParams P) {
    const int f = blockIdx.x;
    if (!((P.cam_mask >> f) & 1u)) return;
    __shared__ float red[12][32];
    float acc[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) acc[k] = 0.0f;
    // this rank's rays of frame f: a contiguous run in either ray order
    const int r_lo = P.order.identity() ? max(P.lo, f * P.pix_per_frame) : P.lo + f * P.order.ppr;
    const int r_hi = P.order.identity() ? min(P.hi, (f + 1) * P.pix_per_frame) : min(P.hi, P.lo + (f + 1) * P.order.ppr);
    for (int i = r_lo + threadIdx.x; i < r_hi; i += blockDim.x) {
        if (P.valid && !P.valid[i]) continue;
        const int64_t id = pool_row(P.idx, P.pool, P.pool_iters, P.it, P.n_total)[P.order.source(i)];
        const float xf = (float)(P.W0 + (int)(id % P.Wc)), yf = (float)(P.H0 + (int)(id / P.Wc));
        const float dir[3] = {(xf - P.cx) / P.fx, P.raydir == 0 ? (xf - P.cy) / P.fy : -(yf - P.cy) / P.fy, -1.0f};
        const float* g = P.d_rays + 6 * (size_t)i;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            acc[9 + r] += g[r];                                   // rays_o = t
#pragma unroll
            for (int c = 0; c < 3; ++c) acc[3 * r + c] += g[3 + r] * dir[c];   // rays_d = R dir
        }
    }
    const int l = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 12; ++k) { const float v = warp_sum(acc[k]); if (l == 0) red[k][w] = v; }
    __syncthreads();
    if (w == 0) {
        const int nw = blockDim.x >> 5;
        float G[12];
#pragma unroll
        for (int k = 0; k < 12; ++k) G[k] = warp_sum(l < nw ? red[k][l] : 0.0f);
        if (l == 0) {
            float gq[4];
            quad2rotation_vjp(P.cams + 8 * f, G, gq);
            float* o = P.g_cams + 8 * f;
            o[0] = gq[0]; o[1] = gq[1]; o[2] = gq[2]; o[3] = gq[3]; o[4] = G[9]; o[5] = G[10]; o[6] = G[11];
        }
    }
}


// ---- keyframe selection by view overlap: Mapper::keyframe_selection_overlap (Mapper.cpp:132-196) ------------------------
struct OverlapParams {
    const float* rays_o; const float* rays_d; const float* gt_depth;   // the `pixels` rays sampled from the current frame
    const float* t_vals;     // linspace(0, 1, n_samples)
    const float* w2c;        // [n_kf][12] inverse keyframe poses, row-major 3x4
    int pixels, n_samples, n_kf;
    int H, W, edge;
    float fx, fy, cx, cy;
    float* percent;          // [n_kf] fraction of the pixels x n_samples vertices that project inside the keyframe
};

// One block per keyframe.  Vertices z = 0.8 d (1 - t) + (d + 0.5) t along each ray (:142-146) are projected with
// K [-x, y, z] (:166-170) and counted when 20 px inside the image and in front of the camera (z < 0, :172-174).
__global__ void k_kf_overlap(OverlapParams P) {
    const int kf = blockIdx.x;
    const float* M = P.w2c + 12 * kf;
    const int nv = P.pixels * P.n_samples;
    int cnt = 0;
    for (int v = threadIdx.x; v < nv; v += blockDim.x) {
        const int r = v / P.n_samples, k = v % P.n_samples;
        const float d = P.gt_depth[r], t = P.t_vals[k];
        const float nearv = __fmul_rn(d, 0.8f), farv = __fadd_rn(d, 0.5f);
        const float z = __fadd_rn(__fmul_rn(nearv, __fsub_rn(1.0f, t)), __fmul_rn(farv, t));
        float p[3], c[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) p[a] = __fadd_rn(P.rays_o[3 * r + a], __fmul_rn(P.rays_d[3 * r + a], z));
#pragma unroll
        for (int a = 0; a < 3; ++a) c[a] = M[4 * a] * p[0] + M[4 * a + 1] * p[1] + M[4 * a + 2] * p[2] + M[4 * a + 3];
        c[0] = -c[0];
        const float zz = c[2] + 1e-5f;
        const float u = (P.fx * c[0] + P.cx * c[2]) / zz, w = (P.fy * c[1] + P.cy * c[2]) / zz;
        const bool in = u < (float)(P.W - P.edge) && u > (float)P.edge && w < (float)(P.H - P.edge) && w > (float)P.edge && zz < 0.0f;
        cnt += in ? 1 : 0;
    }
    __shared__ int red[32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x < 32) {
        int v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (threadIdx.x == 0) P.percent[kf] = (float)v / (float)nv;
    }
}

// Tracking variant: several blocks reduce the rays, the last block to finish turns the 12 sums into d L / d (q, t) and applies
// the Adam step to the 7-vector in place (torch::optim::Adam on one tiny parameter, Tracker.cpp:85,103) -- one launch instead
// of a single-block reduction plus an optimiser launch.
struct TrackStepParams {
    PoseGradParams G;
    float* partial;          // [13]: 12 sums + block counter (as int), zero on entry, left zero on exit
    float* cam; float* m; float* v;   // the 7-vector and its Adam moments
    float* g_out;            // [7] gradient copy for the caller (may alias the gradient arena)
    float beta1, beta2, om_beta1, om_beta2, eps, bc2_sqrt, step;   // step = lr / (1 - beta1^t)
};

__global__ void __launch_bounds__(256) k_track_step(TrackStepParams T) {
    const PoseGradParams& P = T.G;
    __shared__ float red[12][8];
    __shared__ int s_last;
    float acc[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) acc[k] = 0.0f;
    for (int i = P.lo + blockIdx.x * blockDim.x + threadIdx.x; i < P.hi; i += gridDim.x * blockDim.x) {
        if (P.valid && !P.valid[i]) continue;
        const int64_t id = P.idx[i];
        const float xf = (float)(P.W0 + (int)(id % P.Wc)), yf = (float)(P.H0 + (int)(id / P.Wc));
        const float dir[3] = {(xf - P.cx) / P.fx, P.raydir == 0 ? (xf - P.cy) / P.fy : -(yf - P.cy) / P.fy, -1.0f};
        const float* g = P.d_rays + 6 * (size_t)i;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            acc[9 + r] += g[r];
#pragma unroll
            for (int c = 0; c < 3; ++c) acc[3 * r + c] += g[3 + r] * dir[c];
        }
    }
    const int l = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 12; ++k) { const float v = warp_sum(acc[k]); if (l == 0) red[k][w] = v; }
    __syncthreads();
    if (threadIdx.x < 12) {
        float v = 0.0f;
        for (int q = 0; q < (int)(blockDim.x >> 5); ++q) v += red[threadIdx.x][q];
        atomicAdd(T.partial + threadIdx.x, v);
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(reinterpret_cast<int*>(T.partial + 12), 1) == (int)gridDim.x - 1);
    __syncthreads();
    if (!s_last || threadIdx.x != 0) return;
    __threadfence();
    float G[12];
    for (int k = 0; k < 12; ++k) { G[k] = __ldcg(T.partial + k); T.partial[k] = 0.0f; }
    reinterpret_cast<int*>(T.partial)[12] = 0;
    float gq[7];
    quad2rotation_vjp(T.cam, G, gq);
    gq[4] = G[9]; gq[5] = G[10]; gq[6] = G[11];
    for (int k = 0; k < 7; ++k) {      // same operation order as k_adam / libtorch
        const float gk = gq[k];
        if (T.g_out) T.g_out[k] = gk;
        const float mm = __fadd_rn(__fmul_rn(T.m[k], T.beta1), __fmul_rn(T.om_beta1, gk));
        const float vv = __fadd_rn(__fmul_rn(T.v[k], T.beta2), __fmul_rn(__fmul_rn(T.om_beta2, gk), gk));
        const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(vv), T.bc2_sqrt), T.eps);
        T.m[k] = mm; T.v[k] = vv;
        T.cam[k] = __fadd_rn(T.cam[k], __fdiv_rn(__fmul_rn(-T.step, mm), denom));
    }
}

// ---- frustum voxel mask: Mapper::get_mask_from_c2w (Mapper.cpp:42-130, upstream semantics -- the transliteration truncates
// the bound to int and mixes up the memcpy directions / the sign of z) --------------------------------------
struct FrustumParams {
    const float* depth;      // (H, W) frame
    float w2c[12];           // inverse pose, row-major 3x4
    float cam_o[3];          // camera centre (c2w[:3, 3])
    Bound bnd;
    int H, W, Z, Y, X;
    float fx, fy, cx, cy;
    float* vdepth;           // [Z*Y*X] remapped depth per voxel
    float* stats;            // [0] max remapped depth (int bits)
    uint8_t* mask;           // [Z][Y][X]
};

// torch::linspace(lo, hi, n)[i] in fp32 (symmetric formula)
__device__ __forceinline__ float linspace_at(float lo, float hi, int n, int i) {
    const float step = __fdiv_rn(__fsub_rn(hi, lo), (float)(n - 1));
    return i < n / 2 ? __fadd_rn(lo, __fmul_rn(step, (float)i)) : __fsub_rn(hi, __fmul_rn(step, (float)(n - 1 - i)));
}

__device__ __forceinline__ void frustum_project(const FrustumParams& P, int v, float (&pw)[3], float& u, float& vv, float& z) {
    const int x = v % P.X, y = (v / P.X) % P.Y, zz = v / (P.X * P.Y);
    pw[0] = linspace_at(P.bnd.lo[0], P.bnd.hi[0], P.X, x);
    pw[1] = linspace_at(P.bnd.lo[1], P.bnd.hi[1], P.Y, y);
    pw[2] = linspace_at(P.bnd.lo[2], P.bnd.hi[2], P.Z, zz);
    float c[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) c[r] = P.w2c[4 * r] * pw[0] + P.w2c[4 * r + 1] * pw[1] + P.w2c[4 * r + 2] * pw[2] + P.w2c[4 * r + 3];
    c[0] = -c[0];                                            // cam_cord[:, 0] *= -1
    z = c[2] + 1e-5f;                                        // z = uv[:, -1:] + 1e-5
    u = (P.fx * c[0] + P.cx * c[2]) / z;                     // uv = K @ cam_cord, uv[:, :2] / z
    vv = (P.fy * c[1] + P.cy * c[2]) / z;
}

// cv::remap(depth, u, v, INTER_LINEAR, BORDER_CONSTANT 0) for one point: coordinates quantised to 1/32 pixel like OpenCV
__device__ __forceinline__ float remap_bilinear(const FrustumParams& P, float u, float v) {
    if (!(fabsf(u) < 1e6f) || !(fabsf(v) < 1e6f)) return 0.0f;
    const int sx = __float2int_rn(u * 32.0f), sy = __float2int_rn(v * 32.0f);
    const int ix = sx >> 5, iy = sy >> 5;
    const float a = (float)(sx & 31) * (1.0f / 32.0f), b = (float)(sy & 31) * (1.0f / 32.0f);
    auto at = [&](int yy, int xx) { return (xx >= 0 && xx < P.W && yy >= 0 && yy < P.H) ? P.depth[(size_t)yy * P.W + xx] : 0.0f; };
    const float w0 = (1.0f - b) * (1.0f - a), w1 = (1.0f - b) * a, w2 = b * (1.0f - a), w3 = b * a;
    return at(iy, ix) * w0 + at(iy, ix + 1) * w1 + at(iy + 1, ix) * w2 + at(iy + 1, ix + 1) * w3;
}

__global__ void k_frustum_depth(FrustumParams P) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    float d = 0.0f;
    if (v < P.Z * P.Y * P.X) {
        float pw[3], u, vv, z;
        frustum_project(P, v, pw, u, vv, z);
        d = remap_bilinear(P, u, vv);
        P.vdepth[v] = d;
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) d = fmaxf(d, __shfl_xor_sync(0xffffffffu, d, s));
    if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<int*>(P.stats), __float_as_int(fmaxf(d, 0.0f)));
}

__global__ void k_frustum_mask(FrustumParams P) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= P.Z * P.Y * P.X) return;
    float pw[3], u, vv, z;
    frustum_project(P, v, pw, u, vv, z);
    float d = P.vdepth[v];
    if (d == 0.0f) d = P.stats[0];                           // depths[zero_mask] = max(depths)
    bool m = (u < (float)P.W) && (u > 0.0f) && (vv < (float)P.H) && (vv > 0.0f);   // edge = 0
    m = m && (0.0f <= -z) && (-z <= d + 0.5f);               // depth test
    const float dx = pw[0] - P.cam_o[0], dy = pw[1] - P.cam_o[1], dz = pw[2] - P.cam_o[2];
    m = m || (dx * dx + dy * dy + dz * dz < 0.25f);          // features near the camera centre
    P.mask[v] = m ? 1 : 0;
}

}  // namespace nsb
