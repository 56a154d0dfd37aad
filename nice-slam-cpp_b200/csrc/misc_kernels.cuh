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

// One contiguous run of the parameter arena that shares a learning rate (an Adam "param group" of
// Mapper.cpp:330: decoders, coarse, middle, fine, color grids, camera tensors).
struct AdamSegment {
    int begin, end;          // float offsets into the arena, multiples of 4
    float step;              // lr / (1 - beta1^t), computed in double on the host like libtorch's step_size
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
    float bc2_sqrt;          // (float)sqrt(1 - beta2^t), double arithmetic on the host
    float grad_scale;        // 1 (single GPU) -- kept for mean-style reductions
};

// torch::optim::Adam::step (libtorch defaults, no amsgrad / weight decay) + zero_grad, one launch, one float4 per thread
// over the concatenation of all segments (every load of the 90 MB pass is in flight at once):
//   m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g g;  p -= (lr / bc1) * m / (sqrt(v)/sqrt(bc2) + eps);  g = 0
__global__ void k_adam(AdamParams P) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= P.cum4[P.n_seg]) return;
    int s = 0;
#pragma unroll
    for (int k = 1; k < ADAM_MAX_SEG; ++k) if (k < P.n_seg && tid >= P.cum4[k]) s = k;
    const AdamSegment sg = P.seg[s];
    const int i = sg.begin / 4 + (tid - P.cum4[s]);
    const float nstep = -sg.step;
    float4* gp = reinterpret_cast<float4*>(P.grad) + i;
    float4 g = *gp;
    *gp = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!sg.active) return;
    if (sg.mask && !sg.mask[(i * 4 - sg.begin) / CDIM]) return;
    float4 m = reinterpret_cast<float4*>(P.m)[i], v = reinterpret_cast<float4*>(P.v)[i], p = reinterpret_cast<float4*>(P.param)[i];
    float* gg = reinterpret_cast<float*>(&g); float* mm = reinterpret_cast<float*>(&m);
    float* vv = reinterpret_cast<float*>(&v); float* pp = reinterpret_cast<float*>(&p);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float gk = gg[k] * P.grad_scale;
        mm[k] = __fadd_rn(__fmul_rn(mm[k], P.beta1), __fmul_rn(P.om_beta1, gk));                   // mul_(b1).add_(g, 1-b1)
        vv[k] = __fadd_rn(__fmul_rn(vv[k], P.beta2), __fmul_rn(__fmul_rn(P.om_beta2, gk), gk));   // mul_(b2).addcmul_(g, g, 1-b2)
        const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(vv[k]), P.bc2_sqrt), P.eps);
        pp[k] = __fadd_rn(pp[k], __fdiv_rn(__fmul_rn(nstep, mm[k]), denom));                      // addcdiv_(m, denom, -step)
    }
    reinterpret_cast<float4*>(P.m)[i] = m; reinterpret_cast<float4*>(P.v)[i] = v; reinterpret_cast<float4*>(P.param)[i] = p;
}

__global__ void k_fill(float* p, float v, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

}  // namespace nsb
