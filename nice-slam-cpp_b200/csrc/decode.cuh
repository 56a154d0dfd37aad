// Fused decoder kernels: trilinear grid gather -> Gaussian Fourier features -> width-32 MLP on
// tensor cores (mma.sync m16n8k8 tf32, optional 3xTF32 split), one 16-sample tile per warp, activations
// chained through registers (the C fragment of one layer is the A fragment of the next), weights resident
// in shared memory.  Replaces NICE::forward / MLP::forward / GaussianFFT::forward / grid_sample
// (NICE.cpp:16-51, MLP.cpp:51-102, GaussianFFT.cpp:10-15) and their autograd backward.
#pragma once
#include "common.cuh"
#include "params.h"

namespace nsb {

// Offsets (floats) of one decoder inside its flat parameter vector (layout documented in include/nsb.h).
struct DecFlat {
    int B, W[5], b[5], Fc[5], bc[5], Wo, bo, total;
    __host__ __device__ static DecFlat make(int C, int O) {
        DecFlat f;
        const int K[5] = {EMB, HID, HID, EMB + HID, HID};
        int off = 0;
        f.B = off; off += 3 * EMB;
        for (int i = 0; i < 5; ++i) { f.W[i] = off; off += HID * K[i]; f.b[i] = off; off += HID; }
        for (int i = 0; i < 5; ++i) { f.Fc[i] = off; off += HID * C; f.bc[i] = off; off += HID; }
        f.Wo = off; off += O * HID; f.bo = off; off += O;
        f.total = off;
        return f;
    }
};

// Shared-memory image of one decoder (floats).  Every weight matrix starts on a multiple of 32 floats.
template <int C>
struct DecSmem {
    static constexpr int B = 0;                         // [3][96]
    static constexpr int W0 = B + 3 * EMBP;             // [32][96]  swizzled
    static constexpr int W1 = W0 + HID * EMBP;          // [32][32]
    static constexpr int W2 = W1 + HID * HID;
    static constexpr int W3E = W2 + HID * HID;          // [32][96]  skip, embedding columns
    static constexpr int W3H = W3E + HID * EMBP;        // [32][32]  skip, hidden columns
    static constexpr int W4 = W3H + HID * HID;
    static constexpr int FC = W4 + HID * HID;           // 5 x [32][C], columns permuted to the gather layout
    static constexpr int WO = FC + 5 * HID * C;         // [4][32]
    static constexpr int BIAS = WO + 4 * HID;           // b[5][32]
    static constexpr int BIASC = BIAS + 5 * HID;        // bc[5][32]
    static constexpr int BO = BIASC + 5 * HID;          // bo[4]
    static constexpr int HI_END = BO + 4;
    // the MMA weight matrices [W0, WO) are staged pre-split for the 3xTF32 product: tf32-rounded hi plane in place,
    // residual lo plane LO floats further (saves 6 ALU instructions per B fragment in the MMA loops)
    static constexpr int LO = HI_END - W0;
    static constexpr int TOTAL = HI_END + (WO - W0);
    __host__ __device__ static constexpr int w(int i) { return i == 0 ? W0 : i == 1 ? W1 : i == 2 ? W2 : i == 3 ? W3H : W4; }
};

// Column permutation of the Fc matrices: MMA k-step kk, k-index (2t+r) <-> channel 8t + 2kk + r, so that the
// thread that gathered channels 8t..8t+7 of a voxel corner (two float4) owns exactly the A-fragment elements.
__host__ __device__ __forceinline__ int fc_channel(int ip) {
    const int blk = ip >> 5, i = ip & 31;
    return 32 * blk + 8 * ((i >> 1) & 3) + 2 * (i >> 3) + (i & 1);
}

// Stage one weight matrix W[32 out][cols in] into shared memory as M[rows][ld] (swizzled columns):
//   forward kernels  (T = false): M = W      (rows = out, contraction index of  x W^T  along the columns)
//   backward kernels (T = true):  M = W^T    (rows = in,  contraction index of  g W    along the columns)
// so that BOTH products read their B fragments as conflict-free 64-bit loads of two k-adjacent weights.
// hi plane at `off`: tf32-rounded values.  Second plane LO floats further: the residuals w - hi (3xTF32), or with
// NSB_HYBRID_BF16 the packed bf16 pairs {w, w'} at even columns and {lo, lo'} at odd columns.
template <bool T, typename F>
__device__ __forceinline__ void stage_matrix(float* sm, int off, int LO, int cols, int tid, int nthr, F w) {
    const int nrow = T ? cols : HID, ncol = T ? HID : cols;     // M is [nrow][ncol], ld = ncol
    for (int idx = tid; idx < nrow * ncol / 2; idx += nthr) {
        const int r = idx / (ncol / 2), c0 = 2 * (idx % (ncol / 2));
        const float v0 = T ? w(c0, r) : w(r, c0), v1 = T ? w(c0 + 1, r) : w(r, c0 + 1);
        const float h0 = __uint_as_float(f2tf32(v0)), h1 = __uint_as_float(f2tf32(v1));
        const int p0 = off + r * ncol + (c0 ^ swz(r));
        sm[p0] = h0; sm[p0 + 1] = h1;
#ifndef NSB_HYBRID_BF16
        sm[p0 + LO] = v0 - h0; sm[p0 + 1 + LO] = v1 - h1;
#else
        sm[p0 + LO] = __uint_as_float(pack_bf16(v0, v1));
        sm[p0 + 1 + LO] = __uint_as_float(pack_bf16(v0 - h0, v1 - h1));
#endif
    }
}

// BWD = true stages the transposed matrices the data-gradient products need; the Fc matrices then keep only the 32
// input channels that carry gradient (stride HID*HID instead of HID*C).
template <int C, int O, bool BWD>
__device__ void stage_decoder(float* sm, const float* __restrict__ flat, int tid, int nthr) {
    using L = DecSmem<C>;
    const DecFlat f = DecFlat::make(C, O);
    for (int i = tid; i < 3 * EMBP; i += nthr) {
        const int d = i / EMBP, c = i % EMBP;
        sm[L::B + i] = c < EMB ? flat[f.B + d * EMB + c] : 0.0f;
    }
    stage_matrix<BWD>(sm, L::W0, L::LO, EMBP, tid, nthr, [&](int o, int c) { return c < EMB ? flat[f.W[0] + o * EMB + c] : 0.0f; });
    stage_matrix<BWD>(sm, L::W3E, L::LO, EMBP, tid, nthr, [&](int o, int c) { return c < EMB ? flat[f.W[3] + o * (EMB + HID) + c] : 0.0f; });
    stage_matrix<BWD>(sm, L::W1, L::LO, HID, tid, nthr, [&](int o, int c) { return flat[f.W[1] + o * HID + c]; });
    stage_matrix<BWD>(sm, L::W2, L::LO, HID, tid, nthr, [&](int o, int c) { return flat[f.W[2] + o * HID + c]; });
    stage_matrix<BWD>(sm, L::W4, L::LO, HID, tid, nthr, [&](int o, int c) { return flat[f.W[4] + o * HID + c]; });
    stage_matrix<BWD>(sm, L::W3H, L::LO, HID, tid, nthr, [&](int o, int c) { return flat[f.W[3] + o * (EMB + HID) + EMB + c]; });
    for (int l = 0; l < 5; ++l) {
        if (BWD) stage_matrix<true>(sm, L::FC + l * HID * HID, L::LO, HID, tid, nthr, [&](int o, int ip) { return flat[f.Fc[l] + o * C + fc_channel(ip)]; });
        else stage_matrix<false>(sm, L::FC + l * HID * C, L::LO, C, tid, nthr, [&](int o, int ip) { return flat[f.Fc[l] + o * C + fc_channel(ip)]; });
    }
    for (int i = tid; i < 4 * HID; i += nthr) sm[L::WO + i] = (i / HID) < O ? flat[f.Wo + i] : 0.0f;
    for (int i = tid; i < 5 * HID; i += nthr) {
        sm[L::BIAS + i] = flat[f.b[i / HID] + i % HID];
        sm[L::BIASC + i] = flat[f.bc[i / HID] + i % HID];
    }
    if (tid < 4) sm[L::BO + tid] = tid < O ? flat[f.bo + tid] : 0.0f;
}

// Gather the thread's 8 channels (8t..8t+7) of the trilinear feature for its two samples.
__device__ __forceinline__ void gather8(const GridView& G, const Bound& bnd, const float (&p)[3], int t, float* c) {
    Tri s;
    tri_setup(G, bnd, p, s);
#pragma unroll
    for (int i = 0; i < 8; ++i) c[i] = 0.0f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        int off;
        const float w = tri_corner(G, s, k, off);
        const float4 v0 = ldg4(G.data + off + 8 * t), v1 = ldg4(G.data + off + 8 * t + 4);
        c[0] = fmaf(v0.x, w, c[0]); c[1] = fmaf(v0.y, w, c[1]); c[2] = fmaf(v0.z, w, c[2]); c[3] = fmaf(v0.w, w, c[3]);
        c[4] = fmaf(v1.x, w, c[4]); c[5] = fmaf(v1.y, w, c[5]); c[6] = fmaf(v1.z, w, c[6]); c[7] = fmaf(v1.w, w, c[7]);
    }
}

__device__ __forceinline__ void init_bias(float (&acc)[4][4], const float* __restrict__ b, int t) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float2 v = *reinterpret_cast<const float2*>(b + 8 * j + 2 * t);
        acc[j][0] = v.x; acc[j][1] = v.y; acc[j][2] = v.x; acc[j][3] = v.y;
    }
}
__device__ __forceinline__ void add_bias(float (&acc)[4][4], const float* __restrict__ b, int t) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float2 v = *reinterpret_cast<const float2*>(b + 8 * j + 2 * t);
        acc[j][0] += v.x; acc[j][1] += v.y; acc[j][2] += v.x; acc[j][3] += v.y;
    }
}

// Pre-activation of layer 0 (acc0) and the embedding part of the skip layer (accS), sharing the sines.
template <int C, bool P3, bool STASH>
__device__ __forceinline__ void embed_layers(const float* __restrict__ sm, const float (&p)[2][3], int g, int t,
                                             float (&acc0)[4][4], float (&accS)[4][4], float* st0, float* st1) {
    using L = DecSmem<C>;
    init_bias(acc0, sm + L::BIAS + 0 * HID, t);
    init_bias(accS, sm + L::BIAS + 3 * HID, t);
#pragma unroll 2
    for (int kk = 0; kk < EMBP / 8; ++kk) {
        const int f0 = 8 * kk + 2 * t;
        const float2 B0 = *reinterpret_cast<const float2*>(sm + L::B + f0);
        const float2 B1 = *reinterpret_cast<const float2*>(sm + L::B + EMBP + f0);
        const float2 B2 = *reinterpret_cast<const float2*>(sm + L::B + 2 * EMBP + f0);
        const float e00 = ff_sin(fmaf(p[0][2], B2.x, fmaf(p[0][1], B1.x, p[0][0] * B0.x)));
        const float e01 = ff_sin(fmaf(p[0][2], B2.y, fmaf(p[0][1], B1.y, p[0][0] * B0.y)));
        const float e10 = ff_sin(fmaf(p[1][2], B2.x, fmaf(p[1][1], B1.x, p[1][0] * B0.x)));
        const float e11 = ff_sin(fmaf(p[1][2], B2.y, fmaf(p[1][1], B1.y, p[1][0] * B0.y)));
        if (STASH) {
            *reinterpret_cast<float2*>(st0 + stash::E + f0) = make_float2(e00, e01);
            *reinterpret_cast<float2*>(st1 + stash::E + f0) = make_float2(e10, e11);
        }
        AFrag<P3> a;
        a.set(e00, e10, e01, e11);
        kstep_fwd<P3, 4>(acc0, a, sm + L::W0, EMBP, kk, g, t, L::LO);
        kstep_fwd<P3, 4>(accS, a, sm + L::W3E, EMBP, kk, g, t, L::LO);
    }
}

// h += c . Fc_l^T   (c in the gather layout: c[r][2kk], c[r][2kk+1] are the k-step kk elements).
// For C == 32 the four A fragments of c are split once (ca) and reused by all five layers.
template <int C, bool P3>
__device__ __forceinline__ void add_cterm(const float* __restrict__ sm, int l, const float (&c)[2][C / 4], const AFrag<P3>* ca,
                                          int g, int t, float (&h)[4][4]) {
    using L = DecSmem<C>;
#pragma unroll
    for (int kk = 0; kk < C / 8; ++kk) {
        if (C == 32) {
            kstep_fwd<P3, 4>(h, ca[kk], sm + L::FC + l * HID * C, C, kk, g, t, L::LO);
        } else {
            AFrag<P3> a;
            a.set(c[0][2 * kk], c[1][2 * kk], c[0][2 * kk + 1], c[1][2 * kk + 1]);
            kstep_fwd<P3, 4>(h, a, sm + L::FC + l * HID * C, C, kk, g, t, L::LO);
        }
    }
}

__device__ __forceinline__ uint32_t relu_mask(float (&h)[4][4], const float (&acc)[4][4]) {
    uint32_t m = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const bool pos = acc[j][q] > 0.0f;
            m |= (pos ? 1u : 0u) << (4 * j + q);
            h[j][q] = pos ? acc[j][q] : 0.0f;
        }
    return m;
}

// C-layout tile (rows g / g+8, features 8j+2t, 8j+2t+1) -> 32 floats of two stash rows
__device__ __forceinline__ void stash_tile(float* st0, float* st1, int off, const float (&h)[4][4], int t) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        *reinterpret_cast<float2*>(st0 + off + 8 * j + 2 * t) = make_float2(h[j][0], h[j][1]);
        *reinterpret_cast<float2*>(st1 + off + 8 * j + 2 * t) = make_float2(h[j][2], h[j][3]);
    }
}

// Full decoder forward for one 16-sample tile.  out[r][o]: r = 0 -> row g, r = 1 -> row g+8 (valid in all 4
// lanes of the quad).  masks[i] = relu pattern of layer i (bit 4j+q).  With STASH the embedding and the block
// outputs h_1..h_5 are written to the two samples' stash rows (st0, st1) for the weight-gradient kernel.
template <int C, int O, bool P3, bool STASH>
__device__ __forceinline__ void decoder_forward(const float* __restrict__ sm, const float (&p)[2][3],
                                                const float (&c)[2][C / 4], int g, int t, float (&out)[2][4],
                                                uint32_t (&masks)[5], float (&h)[4][4], float* st0, float* st1) {
    using L = DecSmem<C>;
    float acc[4][4], accS[4][4];
    AFrag<P3> ca[C == 32 ? 4 : 1];
    if (C == 32) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) ca[kk].set(c[0][2 * kk], c[1][2 * kk], c[0][2 * kk + 1], c[1][2 * kk + 1]);
    }
    embed_layers<C, P3, STASH>(sm, p, g, t, acc, accS, st0, st1);
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        if (i > 0) {
            if (i == 3) {
#pragma unroll
                for (int j = 0; j < 4; ++j)
#pragma unroll
                    for (int q = 0; q < 4; ++q) acc[j][q] = accS[j][q];
            } else {
                init_bias(acc, sm + L::BIAS + i * HID, t);
            }
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                AFrag<P3> a;
                afrag_from_c<P3>(a, h[kk]);
                kstep_fwd<P3, 4>(acc, a, sm + L::w(i), HID, kk, g, t, L::LO);
            }
        }
        masks[i] = relu_mask(h, acc);
        add_bias(h, sm + L::BIASC + i * HID, t);
        add_cterm<C, P3>(sm, i, c, ca, g, t, h);
        if (STASH) stash_tile(st0, st1, stash::H + HID * i, h, t);
    }
    constexpr int NO = O == 4 ? 3 : 1;   // the colour decoder's 4th output is overwritten (NICE.cpp:49)
#pragma unroll
    for (int o = 0; o < NO; ++o) {
        float s0 = 0.0f, s1 = 0.0f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 w = *reinterpret_cast<const float2*>(sm + L::WO + o * HID + 8 * j + 2 * t);
            s0 = fmaf(h[j][0], w.x, s0); s0 = fmaf(h[j][1], w.y, s0);
            s1 = fmaf(h[j][2], w.x, s1); s1 = fmaf(h[j][3], w.y, s1);
        }
        out[0][o] = quad_sum(s0) + sm[L::BO + o];
        out[1][o] = quad_sum(s1) + sm[L::BO + o];
    }
}

// ---- coarse decoder (MLP_no_xyz, MLP.cpp:104-181): no embedding, skip = cat(c, h) ----------------------------
struct CoarseFlat {
    int W[5], b[5], Wo, bo, total;
    __host__ __device__ static CoarseFlat make() {
        CoarseFlat f; const int K[5] = {CDIM, HID, HID, CDIM + HID, HID};
        int off = 0;
        for (int i = 0; i < 5; ++i) { f.W[i] = off; off += HID * K[i]; f.b[i] = off; off += HID; }
        f.Wo = off; off += HID; f.bo = off; off += 1; f.total = off;
        return f;
    }
};
struct CoarseSmem {
    static constexpr int W0 = 0;                 // [32][32], columns in the gather layout
    static constexpr int W1 = W0 + 1024, W2 = W1 + 1024;
    static constexpr int W3C = W2 + 1024;        // skip, c columns (gather layout)
    static constexpr int W3H = W3C + 1024, W4 = W3H + 1024;
    static constexpr int WO = W4 + 1024;         // [32]
    static constexpr int BIAS = WO + 32;         // b[5][32]
    static constexpr int BO = BIAS + 160;
    static constexpr int TOTAL = BO + 4;
};
__device__ inline void stage_coarse(float* sm, const float* __restrict__ flat, int tid, int nthr) {
    using L = CoarseSmem;
    const CoarseFlat f = CoarseFlat::make();
    for (int i = tid; i < HID * HID; i += nthr) {
        const int o = i / HID, c = i % HID, d = o * HID + (c ^ swz(o));
        sm[L::W0 + d] = flat[f.W[0] + o * CDIM + fc_channel(c)];
        sm[L::W1 + d] = flat[f.W[1] + i];
        sm[L::W2 + d] = flat[f.W[2] + i];
        sm[L::W3C + d] = flat[f.W[3] + o * (CDIM + HID) + fc_channel(c)];
        sm[L::W3H + d] = flat[f.W[3] + o * (CDIM + HID) + CDIM + c];
        sm[L::W4 + d] = flat[f.W[4] + i];
    }
    for (int i = tid; i < HID; i += nthr) sm[L::WO + i] = flat[f.Wo + i];
    for (int i = tid; i < 5 * HID; i += nthr) sm[L::BIAS + i] = flat[f.b[i / HID] + i % HID];
    if (tid == 0) sm[L::BO] = flat[f.bo];
}
template <bool P3>
__device__ __forceinline__ void coarse_forward(const float* __restrict__ sm, const float (&c)[2][8], int g, int t, float (&out)[2]) {
    using L = CoarseSmem;
    float acc[4][4], h[4][4];
    const int wofs[5] = {L::W0, L::W1, L::W2, L::W3H, L::W4};
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        init_bias(acc, sm + L::BIAS + i * HID, t);
        if (i == 0 || i == 3) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                AFrag<P3> a;
                a.set(c[0][2 * kk], c[1][2 * kk], c[0][2 * kk + 1], c[1][2 * kk + 1]);
                kstep_fwd<P3, 4>(acc, a, sm + (i == 0 ? L::W0 : L::W3C), HID, kk, g, t);
            }
        }
        if (i > 0) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                AFrag<P3> a;
                afrag_from_c<P3>(a, h[kk]);
                kstep_fwd<P3, 4>(acc, a, sm + wofs[i], HID, kk, g, t);
            }
        }
        relu_mask(h, acc);
    }
    float s0 = 0.0f, s1 = 0.0f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float2 w = *reinterpret_cast<const float2*>(sm + L::WO + 8 * j + 2 * t);
        s0 = fmaf(h[j][0], w.x, s0); s0 = fmaf(h[j][1], w.y, s0);
        s1 = fmaf(h[j][2], w.x, s1); s1 = fmaf(h[j][3], w.y, s1);
    }
    out[0] = quad_sum(s0) + sm[L::BO];
    out[1] = quad_sum(s1) + sm[L::BO];
}

}  // namespace nsb
