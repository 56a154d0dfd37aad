// Shared device helpers for the sm_100a kernels of libnsb.so.
#pragma once
#include <cuda_fp16.h>
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
    float inv_len[3];             // 1 / len: the normalisation multiplies by it (<= 1 ulp from the reference's division; the trilinear
                                  // sample is continuous in the coordinate, so this is far inside the 1e-4 tolerance)
};

// One feature grid, channel-last [Z][Y][X][32] fp32: a voxel corner is one 128-byte line.
struct GridView {
    const float* data;
    float* grad;       // same layout, may be null
    int Z, Y, X;
};

// Per-iteration device state of the mapping loop.  Every scalar that changes from one joint iteration to the next (which slot of
// the statistics / loss ring, which row of the resident pixel-index pool, the Adam step count, the peer-barrier epoch) is READ BY
// THE KERNELS from this block instead of being passed as a launch parameter, so that one captured CUDA graph per iteration
// variant (geometry / colour / bundle adjustment) is replayed unchanged for every iteration of Mapper.cpp:331-465: the host issues
// one cudaGraphLaunch per iteration.  The last block of the optimiser kernel advances it.
//   state[0] = iterations completed since nsb_mapping_begin, state[1] = index-pool cursor, state[2] = blocks-done counter
struct IterRef {
    int* state;          // nullptr: no iteration state (render / tracking entry points): slot 0
    int ring;            // slots of the statistics ring
};
__device__ __forceinline__ int iter_step(const IterRef& r) { return r.state ? r.state[0] : 0; }
__device__ __forceinline__ int iter_slot(const IterRef& r) { return r.state ? (r.state[0] % r.ring) : 0; }
__device__ __forceinline__ void zero_tile_counters(unsigned long long* ctr) {   // by one thread of a kernel that precedes the decoder launch
    if (ctr) { ctr[0] = 0ull; ctr[1] = 0ull; ctr[2] = 0ull; ctr[3] = 0ull; }
}

// ---- tensor-core helpers: mma.sync m16n8k16 f16 with fp32 accumulation, fp16 two-way split (fp32-grade accuracy) ----
// Every fp32 operand x is carried as the pair  hi = fp16(x),  lo = fp16(x - hi):  |hi + lo - x| <= max(2^-22 |x|, 2^-25)
// -- 22 significant bits while lo is a normal fp16 (|x| >= 0.125), and below that the half-spacing of the fp16 subnormal
// grid, 3e-8 ABSOLUTE, far under the fp32 rounding noise of the O(0.1..10) sums these operands enter
// (tests/test_split_numerics.py).  A product is three tensor-core instructions  a_lo.b_hi + a_hi.b_lo + a_hi.b_hi  accumulated in fp32 --
// the same error (~2^-22 per product, measured 3-4e-7 on the decoder outputs against fp64) as the 3xTF32 split this
// replaces, at half the instruction count: one m16n8k16 covers 16 contraction indices where the tf32 m16n8k8 covers 8,
// and both issue at the same rate on sm_100a (tools/microbench).
// f2tf32 is kept for the tcgen05 kernel (kind::tf32), which splits into tf32 hi/lo planes.
__device__ __forceinline__ uint32_t f2tf32(float x) { return (__float_as_uint(x) + 0x1000u) & 0xffffe000u; }

// two fp32 -> packed f16x2, first argument in the low half (= the lower k index of the MMA fragment)
__device__ __forceinline__ uint32_t pack_f16(float lo_half, float hi_half) {
    const __half2 v = __floats2half2_rn(lo_half, hi_half);
    return *reinterpret_cast<const uint32_t*>(&v);
}
// (x0, x1) -> hi = f16x2(x0, x1), lo = f16x2(x0 - hi0, x1 - hi1): four instructions per pair.  The residual uses sm_100's
// mixed-precision add (add.f32.f16 -> FHADD, which reads one half of the packed register directly), so the fp16 -> fp32 unpack of
// the older form (two more instructions per pair) disappears; the result is bit-identical (x - hi is exact in fp32 either way).
__device__ __forceinline__ void split_f16(float x0, float x1, uint32_t& hi, uint32_t& lo) {
    float r0, r1;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(x1), "f"(x0));
    asm("{\n\t.reg .b16 l, h, nl, nh;\n\tmov.b32 {l, h}, %2;\n\tneg.f16 nl, l;\n\tneg.f16 nh, h;\n\t"
        "add.rn.f32.f16 %0, nl, %3;\n\tadd.rn.f32.f16 %1, nh, %4;\n\t}" : "=f"(r0), "=f"(r1) : "r"(hi), "f"(x0), "f"(x1));
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(r1), "f"(r0));
}
__device__ __forceinline__ void mma_f16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// A operand of one k-step (16 contraction indices) for the 16 rows of the tile, pre-split.  The thread (g, t) holds, for
// rows g and g+8, the four k slots 2t, 2t+1, 2t+8, 2t+9 of the instruction; `set` takes them in that order.
template <bool P3>
struct AFrag {
    uint32_t hi[4], lo[4];
    __device__ __forceinline__ void set(float r0s0, float r0s1, float r0s2, float r0s3, float r1s0, float r1s1, float r1s2, float r1s3) {
        if (P3) {
            split_f16(r0s0, r0s1, hi[0], lo[0]); split_f16(r1s0, r1s1, hi[1], lo[1]);
            split_f16(r0s2, r0s3, hi[2], lo[2]); split_f16(r1s2, r1s3, hi[3], lo[3]);
        } else {
            hi[0] = pack_f16(r0s0, r0s1); hi[1] = pack_f16(r1s0, r1s1); hi[2] = pack_f16(r0s2, r0s3); hi[3] = pack_f16(r1s2, r1s3);
        }
    }
};

// acc += A . B for a B fragment given as four fp32 pairs: (k slots 2t, 2t+1) and (2t+8, 2t+9) of column g, split on the fly
template <bool P3>
__device__ __forceinline__ void mma_acc(float (&d)[4], const AFrag<P3>& a, float w0, float w1, float w2, float w3) {
    if (P3) {
        uint32_t bh0, bl0, bh1, bl1;
        split_f16(w0, w1, bh0, bl0); split_f16(w2, w3, bh1, bl1);
        mma_f16(d, a.lo, bh0, bh1);   // small terms first
        mma_f16(d, a.hi, bl0, bl1);
        mma_f16(d, a.hi, bh0, bh1);
    } else {
        mma_f16(d, a.hi, pack_f16(w0, w1), pack_f16(w2, w3));
    }
}

// Weight matrices live in shared memory pre-split as M[n][K] (n = output index of the product, K = contraction length,
// a multiple of 16), one row = K 32-bit words + 16 words of padding: k-step kk of row n is the 16-word block
//     word 16 kk + 4 t + {0, 1, 2, 3} = { hi(slots 2t,2t+1), hi(slots 2t+8,2t+9), lo(2t,2t+1), lo(2t+8,2t+9) }
// so a lane fetches its whole B fragment (both planes) with ONE 128-bit load, and the row stride K + 16 = 16 (mod 32)
// words makes the quarter-warp (rows g, g+1 x four t) hit 32 distinct banks.  "Position" pos = 16 kk + 4 t + i of a row
// therefore holds k slot (2t, 2t+1, 2t+8, 2t+9)[i] of k-step kk; which input feature sits at a position is the staging
// code's choice (identity for computed operands, perm16 for operands chained from accumulator fragments).
__host__ __device__ constexpr int wstride(int K) { return K + 16; }
// feature held at position pos when the A operand is an accumulator tile pair (tiles 2kk, 2kk+1) reused in place
__host__ __device__ __forceinline__ int perm16(int pos) {
    const int q = pos & 15, t = q >> 2, i = q & 3;
    return (pos & ~15) + 2 * t + (i & 1) + 8 * (i >> 1);
}

// acc[j] += A(kk) . M[8j+g][k-step kk]^T for the NJ output tiles.
template <bool P3, int NJ>
__device__ __forceinline__ void kstep_fwd(float (&acc)[NJ][4], const AFrag<P3>& a, const uint32_t* __restrict__ M, int stride, int kk, int g, int t) {
    const uint32_t* wp = M + g * stride + 16 * kk + 4 * t;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
        const uint4 b = *reinterpret_cast<const uint4*>(wp + 8 * j * stride);
        if (P3) {
            mma_f16(acc[j], a.lo, b.x, b.y);
            mma_f16(acc[j], a.hi, b.z, b.w);
        }
        mma_f16(acc[j], a.hi, b.x, b.y);
    }
}

// Two adjacent accumulator tiles (rows g, g+8; features 16kk+2t,+1 and 16kk+8+2t,+1) reused as the A operand of
// k-step kk: the fragment layouts coincide, no data movement.
template <bool P3>
__device__ __forceinline__ void afrag_from_c(AFrag<P3>& a, const float (&c0)[4], const float (&c1)[4]) {
    a.set(c0[0], c0[1], c1[0], c1[1], c0[2], c0[3], c1[2], c1[3]);
}

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
        float pn = __fsub_rn(__fmul_rn(__fmul_rn(__fsub_rn(p[a], B.lo[a]), B.inv_len[a]), 2.0f), 1.0f);
        // grid_sampler_unnormalize(align_corners): ((x + 1) / 2) * (size - 1)
        const float sm1 = (float)(dim[a] - 1);
        float ix = __fmul_rn(__fmul_rn(__fadd_rn(pn, 1.0f), 0.5f), sm1);
        const bool clipped = !(ix > 0.0f && ix < sm1);
        ix = fminf(sm1, fmaxf(ix, 0.0f));
        const float f = floorf(ix);
        const int i = (int)f;
        s.i0[a] = i;
        s.i1[a] = min(i + 1, dim[a] - 1);
        s.w1[a] = __fsub_rn(ix, f);
        s.w0[a] = __fsub_rn(__fadd_rn(f, 1.0f), ix);
        s.gm[a] = clipped ? 0.0f : __fmul_rn(sm1, B.inv_len[a]);
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

// the same with the voxel INDEX (z Y + y) X + x instead of the float offset: the caller forms the address as
// base + (unsigned) index * 128 bytes, one IMAD.WIDE.U32 instead of shift + sign extension + 64-bit LEA pair
__device__ __forceinline__ float tri_corner_idx(const GridView& G, const Tri& s, int k, unsigned& idx) {
    const int dx = k & 1, dy = (k >> 1) & 1, dz = (k >> 2) & 1;
    const int x = dx ? s.i1[0] : s.i0[0], y = dy ? s.i1[1] : s.i0[1], z = dz ? s.i1[2] : s.i0[2];
    idx = (unsigned)((z * G.Y + y) * G.X + x);
    const float wx = dx ? s.w1[0] : s.w0[0], wy = dy ? s.w1[1] : s.w0[1], wz = dz ? s.w1[2] : s.w0[2];
    return __fmul_rn(__fmul_rn(wx, wy), wz);
}

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
// 256-bit read-only load (sm_100): the four lanes of a quad fetch one whole 128-byte voxel line in a single L1 wavefront
__device__ __forceinline__ void ldg8(const float* p, float (&v)[8]) {
    asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "l"(p));
}

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
