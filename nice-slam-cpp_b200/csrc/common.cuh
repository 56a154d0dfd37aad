// Shared device helpers for the sm_100a kernels of libnsb.so.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace nsb {

constexpr int EMB = 93;       // Fourier embedding width (MLP.cpp:21)
constexpr int EMBP = 96;      // padded to a multiple of the MMA k-step
constexpr int HID = 32;       // hidden width (main.cpp:29)
constexpr int CDIM = 32;      // grid channels (nice_slam.yaml model.c_dim)
constexpr int TILE = 16;      // samples per warp tile (MMA m16)

// Scene bound and derived constants, all in fp32 exactly as the reference computes them.
struct Bound {
    float lo[3], hi[3], len[3];   // len = hi - lo in fp32 (utils.h:135-137)
};

// One feature grid, channel-last [Z][Y][X][32] fp32: a voxel corner is one 128-byte line.
struct GridView {
    const float* data;
    float* grad;       // same layout, may be null
    int Z, Y, X;
};

// ---- tensor-core helpers: mma.sync m16n8k8 tf32, optional 3xTF32 split (fp32-grade accuracy) ------------
// fp32 -> tf32 with round-to-nearest (ties away), as two integer ops: on sm_100a `cvt.rna.tf32.f32` expands to a
// ~5-instruction FSETP/SEL/LOP3 sequence (NaN/Inf handling) that dominated the issue slots of the MMA loops.
// The tensor core reads only the upper 19 bits of a tf32 operand, so the low part of the split is passed as raw
// fp32 bits (truncation of an already 2^-11-scaled residual: ~2^-22 relative).
__device__ __forceinline__ uint32_t f2tf32(float x) { return (__float_as_uint(x) + 0x1000u) & 0xffffe000u; }
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
    hi = f2tf32(x);
    lo = __float_as_uint(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                         uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

__device__ __forceinline__ void mma_bf16(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// two fp32 -> packed bf16x2, first argument in the low half (= the lower k index of the MMA fragment)
__device__ __forceinline__ uint32_t pack_bf16(float lo_half, float hi_half) {
    const __nv_bfloat162 v = __floats2bfloat162_rn(lo_half, hi_half);
    return *reinterpret_cast<const uint32_t*>(&v);
}

// Default: the three-instruction 3xTF32 split (a_lo.b_hi + a_hi.b_lo + a_hi.b_hi), error ~2^-22 per product.
// Build variant NSB_HYBRID_BF16 (libnsb_hybrid.so): two instructions per k-step of 8 (measured on B200: a bf16 m16n8k16
// issues at the same rate as a tf32 m16n8k8):
//   main term   a_hi . b_hi                       one tf32 m16n8k8
//   cross terms a_lo . b_hi  +  a_hi . b_lo       one bf16 m16n8k16: k slots 0..7 carry (a_lo, b_hi), slots 8..15 (a_hi, b_lo)
// The cross terms are 2^-11 of the product, so bf16 operands (8 bits) leave a ~2^-20 relative error, the same order as
// the 3xTF32 split's dropped a_lo.b_lo term x4.  Measured: forward 0.40 -> 0.345 ms and all mapping parity checks still at
// 2-4e-6, but the ill-conditioned tracking pose gradient (fp32-vs-fp64 oracle noise 8e-5) degrades from 5e-6 to 6e-3,
// above the 1e-3 tolerance -- hence not the default.
//
// A operand of one k-step (8 features) for the 16 rows of the tile, pre-split.
template <bool P3>
struct AFrag {
    uint32_t hi[4];
#ifndef NSB_HYBRID_BF16
    uint32_t lo[4];
#else
    uint32_t x[4];   // bf16 fragment: x0/x1 = lo parts of rows g / g+8 (k slots 2t,2t+1), x2/x3 = bf16(a) (slots 2t+8, 2t+9)
#endif
    // a0 = (row g, feature 2t), a1 = (row g+8, 2t), a2 = (row g, 2t+1), a3 = (row g+8, 2t+1)
    __device__ __forceinline__ void set(float a0, float a1, float a2, float a3) {
        hi[0] = f2tf32(a0); hi[1] = f2tf32(a1); hi[2] = f2tf32(a2); hi[3] = f2tf32(a3);
        if (P3) {
#ifndef NSB_HYBRID_BF16
            lo[0] = __float_as_uint(a0 - __uint_as_float(hi[0])); lo[1] = __float_as_uint(a1 - __uint_as_float(hi[1]));
            lo[2] = __float_as_uint(a2 - __uint_as_float(hi[2])); lo[3] = __float_as_uint(a3 - __uint_as_float(hi[3]));
#else
            x[0] = pack_bf16(a0 - __uint_as_float(hi[0]), a2 - __uint_as_float(hi[2]));
            x[1] = pack_bf16(a1 - __uint_as_float(hi[1]), a3 - __uint_as_float(hi[3]));
            x[2] = pack_bf16(a0, a2);
            x[3] = pack_bf16(a1, a3);
#endif
        }
    }
};

// B fragment given as two fp32 weights (k = 2t and 2t+1 of the k-step, column g): split on the fly.
template <bool P3>
__device__ __forceinline__ void mma_acc(float (&d)[4], const AFrag<P3>& a, float w0, float w1) {
    const uint32_t bh0 = f2tf32(w0), bh1 = f2tf32(w1);
    if (P3) {
#ifndef NSB_HYBRID_BF16
        const uint32_t bl0 = __float_as_uint(w0 - __uint_as_float(bh0)), bl1 = __float_as_uint(w1 - __uint_as_float(bh1));
        mma_tf32(d, a.lo[0], a.lo[1], a.lo[2], a.lo[3], bh0, bh1);   // small terms first
        mma_tf32(d, a.hi[0], a.hi[1], a.hi[2], a.hi[3], bl0, bl1);
#else
        mma_bf16(d, a.x[0], a.x[1], a.x[2], a.x[3], pack_bf16(w0, w1), pack_bf16(w0 - __uint_as_float(bh0), w1 - __uint_as_float(bh1)));
#endif
    }
    mma_tf32(d, a.hi[0], a.hi[1], a.hi[2], a.hi[3], bh0, bh1);
}

// Same with a B operand that was split when the weights were staged: (h0, h1) tf32 weights from the hi plane and
// (x0, x1) from the second plane -- the residuals (pure 3xTF32) or the packed bf16 pairs {w, w'} / {lo, lo'} (hybrid).
template <bool P3>
__device__ __forceinline__ void mma_acc_ps(float (&d)[4], const AFrag<P3>& a, uint32_t h0, uint32_t h1, uint32_t x0, uint32_t x1) {
    if (P3) {
#ifndef NSB_HYBRID_BF16
        mma_tf32(d, a.lo[0], a.lo[1], a.lo[2], a.lo[3], h0, h1);
        mma_tf32(d, a.hi[0], a.hi[1], a.hi[2], a.hi[3], x0, x1);
#else
        mma_bf16(d, a.x[0], a.x[1], a.x[2], a.x[3], x0, x1);
#endif
    }
    mma_tf32(d, a.hi[0], a.hi[1], a.hi[2], a.hi[3], h0, h1);
}

// Weight matrices W[out][in] live in shared memory with row stride ld (a multiple of 32 floats) and the
// column index XOR-swizzled by the row:  col' = col ^ swz(row),  swz(row) = ((row ^ (row>>1)) & 3) << 3.
// That makes the B-fragment access  M[8j+g][8kk+2t .. +1]  (one 64-bit load per lane, rows vary with g) bank-conflict
// free; the backward kernels stage the transposed matrices so that they use the very same access.
__host__ __device__ __forceinline__ int swz(int row) { return ((row ^ (row >> 1)) & 3) << 3; }

// acc[j] += A(kk) * W[8j+g][8kk+2t..]^T for the NJ output tiles (forward: out = x W^T).
// LO != 0: the matrix was staged pre-split, hi plane at W (tf32-rounded), lo plane at W + LO.
template <bool P3, int NJ>
__device__ __forceinline__ void kstep_fwd(float (&acc)[NJ][4], const AFrag<P3>& a, const float* __restrict__ W,
                                          int ld, int kk, int g, int t, int LO = 0) {
    const int col = (8 * kk + 2 * t) ^ swz(g);   // swz(8j+g) == swz(g)
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
        const float* wp = W + (8 * j + g) * ld + col;
        const float2 w = *reinterpret_cast<const float2*>(wp);
        if (LO) {
            float2 wl = make_float2(0.f, 0.f);
            if (P3) wl = *reinterpret_cast<const float2*>(wp + LO);
            mma_acc_ps<P3>(acc[j], a, __float_as_uint(w.x), __float_as_uint(w.y), __float_as_uint(wl.x), __float_as_uint(wl.y));
        } else {
            mma_acc<P3>(acc[j], a, w.x, w.y);
        }
    }
}

// C-fragment (rows g, g+8; cols 2t, 2t+1 of tile kk) reused as the A operand of k-step kk: the k index of
// the MMA is permuted (k=t <-> feature 2t, k=t+4 <-> feature 2t+1), which the B fragments above match.
template <bool P3>
__device__ __forceinline__ void afrag_from_c(AFrag<P3>& a, const float (&c)[4]) { a.set(c[0], c[2], c[1], c[3]); }

// ---- trilinear sampling, identical arithmetic to ATen grid_sampler_3d (bilinear, border, align_corners) ----
struct Tri {
    int i0[3], i1[3];     // x,y,z corner indices (i1 clamped)
    float w0[3], w1[3];   // weights of i0 / i1 per axis
    float gm[3];          // d(index)/d(p) per axis, 0 where the border clamp is active
};

__device__ __forceinline__ void tri_setup(const GridView& G, const Bound& B, const float (&p)[3], Tri& s) {
    const int dim[3] = {G.X, G.Y, G.Z};
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        // utils.h:135-137: ((p - lo) / (hi - lo)) * 2 - 1, each op rounded separately
        float pn = __fsub_rn(__fmul_rn(__fdiv_rn(__fsub_rn(p[a], B.lo[a]), B.len[a]), 2.0f), 1.0f);
        // grid_sampler_unnormalize(align_corners): ((x + 1) / 2) * (size - 1)
        const float sm1 = (float)(dim[a] - 1);
        float ix = __fmul_rn(__fdiv_rn(__fadd_rn(pn, 1.0f), 2.0f), sm1);
        const bool clipped = !(ix > 0.0f && ix < sm1);
        ix = fminf(sm1, fmaxf(ix, 0.0f));
        const float f = floorf(ix);
        const int i = (int)f;
        s.i0[a] = i;
        s.i1[a] = min(i + 1, dim[a] - 1);
        s.w1[a] = __fsub_rn(ix, f);
        s.w0[a] = __fsub_rn(__fadd_rn(f, 1.0f), ix);
        s.gm[a] = clipped ? 0.0f : __fdiv_rn(sm1, B.len[a]);
    }
}

// weight and voxel offset (in floats, channel-last) of corner k in ATen's order tnw,tne,tsw,tse,bnw,bne,bsw,bse
__device__ __forceinline__ float tri_corner(const GridView& G, const Tri& s, int k, int& off) {
    const int dx = k & 1, dy = (k >> 1) & 1, dz = (k >> 2) & 1;
    const int x = dx ? s.i1[0] : s.i0[0], y = dy ? s.i1[1] : s.i0[1], z = dz ? s.i1[2] : s.i0[2];
    off = ((z * G.Y + y) * G.X + x) * CDIM;
    const float wx = dx ? s.w1[0] : s.w0[0], wy = dy ? s.w1[1] : s.w0[1], wz = dz ? s.w1[2] : s.w0[2];
    return __fmul_rn(__fmul_rn(wx, wy), wz);
}

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// Fourier feature sin(x) for |x| up to a few thousand: exact two-term Cody-Waite reduction to [-pi, pi]
// (k < 2^10 so k*2PI_HI is exact in the fma), then the SFU.  Absolute error ~5e-7, far below the
// ~3e-5 rad the fp32 argument p.B itself carries at |x| ~ 300.
__device__ __forceinline__ float reduce_2pi(float x) {
    // round-to-nearest-even via the 1.5 * 2^23 magic add (two FADDs on the FMA pipe) instead of rintf (FRND, on the
    // quarter-rate XU pipe that also serves MUFU.SIN); exact for |x / 2pi| < 2^22
    const float k = __fsub_rn(__fadd_rn(x * 0.15915494309189535f, 12582912.0f), 12582912.0f);
    float r = fmaf(-k, 6.2831854820251465f, x);
    return fmaf(-k, -1.7484555314695172e-07f, r);
}
__device__ __forceinline__ float ff_sin(float x) {
#ifdef NSB_PRECISE_SIN
    return sinf(x);
#else
    return __sinf(reduce_2pi(x));
#endif
}
__device__ __forceinline__ void ff_sincos(float x, float& s, float& c) {
#ifdef NSB_PRECISE_SIN
    sincosf(x, &s, &c);
#else
    const float r = reduce_2pi(x);
    s = __sinf(r);
    c = __cosf(r);
#endif
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float quad_sum(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    return v;
}

}  // namespace nsb
