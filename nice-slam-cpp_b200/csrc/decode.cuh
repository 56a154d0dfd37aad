// Fused decoder kernels: trilinear grid gather -> Gaussian Fourier features -> width-32 MLP on
// tensor cores (mma.sync m16n8k16 f16, fp16 two-way split with fp32 accumulation), one 16-sample tile per warp, activations
// chained through registers (the C fragment of one layer is the A fragment of the next), weights resident
// in shared memory.  Replaces NICE::forward / MLP::forward / GaussianFFT::forward / grid_sample
// (NICE.cpp:16-51, MLP.cpp:51-102, GaussianFFT.cpp:10-15) and their autograd backward.
#pragma once
#include "common.cuh"
#include "params.h"

#ifndef NSB_EMB_UNROLL
#define NSB_EMB_UNROLL 2   // k-steps of the embedding loop unrolled together (tools/sweep_variants.sh: 1 / 2 / 3 / 6 measured)
#endif

namespace nsb {

constexpr int EMB_UNROLL = NSB_EMB_UNROLL;

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

// Shared-memory image of one decoder, in 32-bit words.  The MMA weight matrices are staged pre-split in the layout
// common.cuh documents (row = K + 16 words: fp16 hi/lo pairs of one k-step interleaved in 16-word blocks); the forward
// kernels stage M = W (rows = output features), the backward kernels M = W^T (rows = input features), which for the
// embedding layers has 96 rows -- the W0 / W3E slots are sized for that.
template <int C>
struct DecSmem {
    static constexpr int SH = wstride(HID);             // row stride of a K = 32 matrix
    static constexpr int SE = wstride(EMBP);            // row stride of a K = 96 matrix (forward W0, W3E)
    static constexpr int SC = wstride(C);               // row stride of the forward Fc matrices (K = C)
    static constexpr int B = 0;                         // [3][96] fp32
    static constexpr int W0 = B + 3 * EMBP;             // fwd [32][SE] | bwd W0^T [96][SH]
    static constexpr int W1 = W0 + EMBP * SH;           // [32][SH]
    static constexpr int W2 = W1 + HID * SH;
    static constexpr int W3E = W2 + HID * SH;           // skip layer, embedding columns: fwd [32][SE] | bwd [96][SH]
    static constexpr int W3H = W3E + EMBP * SH;         // skip layer, hidden columns
    static constexpr int W4 = W3H + HID * SH;
    static constexpr int FC = W4 + HID * SH;            // fwd 5 x [32][SC] (K = grid channels in the gather order) | bwd 5 x Fc^T [32][SH]
    static constexpr int FCS = HID * SC;                // forward stride between the five Fc matrices
    static constexpr int WO = FC + 5 * FCS;             // [4][32] fp32
    static constexpr int BIAS = WO + 4 * HID;           // b[5][32]
    static constexpr int BIASC = BIAS + 5 * HID;        // bc[5][32]
    static constexpr int BO = BIASC + 5 * HID;          // bo[4]   (composed image: Wo bc_4 + bo)
    static constexpr int WOC = BO + 4;                  // composed image only: (Wo Fc_4)[4][C] fp32, natural channel order
    static constexpr int TOTAL = (WOC + 4 * C + 3) & ~3;   // multiple of 4 words: the image is copied with 16-byte loads
    __host__ __device__ static constexpr int w(int i) { return i == 0 ? W0 : i == 1 ? W1 : i == 2 ? W2 : i == 3 ? W3H : W4; }
};
static_assert(EMBP * wstride(HID) >= HID * wstride(EMBP), "W0 slot must hold both orientations");

// Grid channel held at position pos of the forward Fc matrices: the thread that gathered channels 8t..8t+7 of a voxel
// corner (one 256-bit load) owns exactly the A-fragment elements -- k-step kk (of the 32-channel block), slot i of lane t
// <-> channel 8t + 4kk + i.
__host__ __device__ __forceinline__ int fc_channel_fwd(int pos) {
    const int blk = pos >> 5, q = pos & 31;
    return 32 * blk + 8 * ((q >> 2) & 3) + 4 * (q >> 4) + (q & 3);
}
// Backward: row n = 8j + 2t + b of Fc^T (an accumulator column) <-> channel 4t + 2(j&1) + b + 16(j>>1), so that lane t ends
// up with the gradient of channels 4t..4t+3 and 16+4t..16+4t+3: each of its two vector reductions then lands, together
// with the other three lanes of the quad, on 64 contiguous bytes (two whole 32-byte sectors) of the voxel's gradient line.
__host__ __device__ __forceinline__ int fc_channel_bwd(int n) {
    const int j = (n >> 3) & 3, t = (n >> 1) & 3, b = n & 1;
    return 4 * t + 2 * (j & 1) + b + 16 * (j >> 1);
}

// Stage M[n][pos] = w(n, pos), n < nrow, pos < K, pre-split into fp16 hi/lo (layout: common.cuh).
template <typename F>
__device__ __forceinline__ void stage_matrix(float* sm, int off, int K, int nrow, int tid, int nthr, F w) {
    __half* smh = reinterpret_cast<__half*>(sm + off);
    const int stride = wstride(K);
    for (int idx = tid; idx < nrow * K; idx += nthr) {
        const int n = idx / K, pos = idx % K;
        const float v = w(n, pos);
        const __half h = __float2half_rn(v);
        const __half l = __float2half_rn(v - __half2float(h));
        const int word = n * stride + (pos & ~15) + 4 * ((pos >> 2) & 3) + ((pos >> 1) & 1);
        smh[2 * word + (pos & 1)] = h;
        smh[2 * (word + 2) + (pos & 1)] = l;
    }
}

// BWD = true stages the transposed matrices the data-gradient products need; the Fc matrices then keep only the 32
// input channels that carry gradient.
template <int C, int O, bool BWD>
__device__ void stage_decoder(float* sm, const float* __restrict__ flat, int tid, int nthr) {
    using L = DecSmem<C>;
    const DecFlat f = DecFlat::make(C, O);
    for (int i = tid; i < 3 * EMBP; i += nthr) {
        const int d = i / EMBP, c = i % EMBP;
        sm[L::B + i] = c < EMB ? flat[f.B + d * EMB + c] : 0.0f;
    }
    if (!BWD) {
        // x W^T: rows = output feature; embedding operands are computed in position order, hidden ones come from accumulators
        stage_matrix(sm, L::W0, EMBP, HID, tid, nthr, [&](int o, int pos) { return pos < EMB ? flat[f.W[0] + o * EMB + pos] : 0.0f; });
        stage_matrix(sm, L::W3E, EMBP, HID, tid, nthr, [&](int o, int pos) { return pos < EMB ? flat[f.W[3] + o * (EMB + HID) + pos] : 0.0f; });
        stage_matrix(sm, L::W1, HID, HID, tid, nthr, [&](int o, int pos) { return flat[f.W[1] + o * HID + perm16(pos)]; });
        stage_matrix(sm, L::W2, HID, HID, tid, nthr, [&](int o, int pos) { return flat[f.W[2] + o * HID + perm16(pos)]; });
        stage_matrix(sm, L::W3H, HID, HID, tid, nthr, [&](int o, int pos) { return flat[f.W[3] + o * (EMB + HID) + EMB + perm16(pos)]; });
        stage_matrix(sm, L::W4, HID, HID, tid, nthr, [&](int o, int pos) { return flat[f.W[4] + o * HID + perm16(pos)]; });
        for (int l = 0; l < 5; ++l)
            stage_matrix(sm, L::FC + l * L::FCS, C, HID, tid, nthr, [&](int o, int pos) { return flat[f.Fc[l] + o * C + fc_channel_fwd(pos)]; });
    } else {
        // g W: rows = input feature, contraction over the output features, whose operand always comes from accumulators
        stage_matrix(sm, L::W0, HID, EMBP, tid, nthr, [&](int n, int pos) { return n < EMB ? flat[f.W[0] + perm16(pos) * EMB + n] : 0.0f; });
        stage_matrix(sm, L::W3E, HID, EMBP, tid, nthr, [&](int n, int pos) { return n < EMB ? flat[f.W[3] + perm16(pos) * (EMB + HID) + n] : 0.0f; });
        stage_matrix(sm, L::W1, HID, HID, tid, nthr, [&](int n, int pos) { return flat[f.W[1] + perm16(pos) * HID + n]; });
        stage_matrix(sm, L::W2, HID, HID, tid, nthr, [&](int n, int pos) { return flat[f.W[2] + perm16(pos) * HID + n]; });
        stage_matrix(sm, L::W3H, HID, HID, tid, nthr, [&](int n, int pos) { return flat[f.W[3] + perm16(pos) * (EMB + HID) + EMB + n]; });
        stage_matrix(sm, L::W4, HID, HID, tid, nthr, [&](int n, int pos) { return flat[f.W[4] + perm16(pos) * HID + n]; });
        // the FC slots of the backward image hold the composed matrices G_l^T: written by stage_backward_G (block-cooperative)
        for (int i = tid; i < 4 * C; i += nthr) {        // (Wo Fc_4)[o][ch]
            const int o = i / C, ch = i % C;
            float acc = 0.0f;
            if (o < O) {
#pragma unroll 8
                for (int m = 0; m < HID; ++m) acc = fmaf(__ldg(flat + f.Wo + o * HID + m), __ldg(flat + f.Fc[4] + m * C + ch), acc);
            }
            sm[L::WOC + i] = acc;
        }
    }
    for (int i = tid; i < 4 * HID; i += nthr) sm[L::WO + i] = (i / HID) < O ? flat[f.Wo + i] : 0.0f;
    for (int i = tid; i < 5 * HID; i += nthr) {
        sm[L::BIAS + i] = flat[f.b[i / HID] + i % HID];
        sm[L::BIASC + i] = flat[f.bc[i / HID] + i % HID];
    }
    if (tid < 4) sm[L::BO + tid] = tid < O ? flat[f.bo + tid] : 0.0f;
}

// Grid-feature gradient through the composed matrices: g_c = g_out (Wo Fc_4) + sum_{l<4} g_u_{l+1} G_l, G_l = W_{l+1} Fc_l (the
// same algebra as the forward), so the product shares its A fragments (the masked g_u) with the main chain.  FC slot l of the
// backward image holds G_l^T restricted to the 32 input channels that carry gradient.  One thread block builds one slot: W_{l+1}
// and Fc_l go through shared memory (this runs after every update of a trained decoder, so it must not be latency-bound).
template <int C, int O>
__device__ void stage_backward_G(float* img, const float* __restrict__ flat, int l) {
    using L = DecSmem<C>;
    __shared__ float sW[HID][HID + 1], sF[HID][HID + 1];
    const DecFlat f = DecFlat::make(C, O);
    const float* Wn = flat + f.W[l + 1] + (l == 2 ? EMB : 0);
    const int ldw = l == 2 ? EMB + HID : HID;
    for (int i = threadIdx.x; i < HID * HID; i += blockDim.x) {
        const int r = i / HID, c = i % HID;
        sW[r][c] = Wn[r * ldw + c];                          // W_{l+1}[o][m]
        sF[r][c] = flat[f.Fc[l] + r * C + c];                // Fc_l[m][channel], first 32 channels
    }
    __syncthreads();
    stage_matrix(img, L::FC + l * HID * L::SH, HID, HID, threadIdx.x, blockDim.x, [&](int n, int pos) {
        const int o = perm16(pos), ch = fc_channel_bwd(n);
        float acc = 0.0f;
#pragma unroll
        for (int m = 0; m < HID; ++m) acc = fmaf(sW[o][m], sF[m][ch], acc);
        return acc;
    });
}

// Copy a decoder's pre-split image (built once per weight update by k_build_wimg, same layout as stage_decoder writes) from
// global into shared memory: coalesced 16-byte loads instead of ~60 scattered fp32 loads + splits per thread and launch.
template <int C>
__device__ __forceinline__ void load_decoder_image(float* sm, const float* __restrict__ img, int tid, int nthr) {
    const uint4* src = reinterpret_cast<const uint4*>(img);
    uint4* dst = reinterpret_cast<uint4*>(sm);
    for (int i = tid; i < DecSmem<C>::TOTAL / 4; i += nthr) dst[i] = __ldg(src + i);
}

// Ticket-based tile scheduler of the decoder kernels: lane 0 draws the next ticket while the current tile is processed.
struct TileQueue {
    unsigned long long* ctr; unsigned long long base; int ntiles; unsigned long long pending;
    __device__ __forceinline__ void init(unsigned long long* c, unsigned long long b, int n, int lane) { ctr = c; base = b; ntiles = n; prefetch(lane); }
    __device__ __forceinline__ void prefetch(int lane) { if (lane == 0) pending = atomicAdd(ctr, 1ull); }
    // tile index of the pending ticket (or -1 when the decoder's tiles are exhausted); draws the one after it
    __device__ __forceinline__ int next(int lane) {
        const unsigned long long tk = __shfl_sync(0xffffffffu, pending, 0);
        const long long tile = (long long)(tk - base);
        if (tile >= ntiles) return -1;
        prefetch(lane);
        return (int)tile;
    }
};

// Composed forward image (k_compose, the algebra of the tcgen05 kernels): with h_{i+1} = relu(a_i) + Fc_i c + bc_i and
// a_{i+1} = W_{i+1} h_{i+1} + b_{i+1} the grid-feature term moves into the NEXT layer's pre-activation,
//     a_{i+1} = W_{i+1} relu(a_i) + G_i c + b'_{i+1},   G_i = W_{i+1} Fc_i,   b'_{i+1} = b_{i+1} + W_{i+1} bc_i,
// so that the c-term MMAs of layer i+1 no longer depend on layer i's relu and overlap its epilogue.  FC slot l holds G_l
// (l = 0..3), BIAS holds b', WOC / BO the grid-feature part and the constant of the output layer.  comp: layout of k_compose.
template <int C, int O>
__device__ void stage_decoder_composed(float* sm, const float* __restrict__ flat, const float* __restrict__ comp, int tid, int nthr) {
    using L = DecSmem<C>;
    const DecFlat f = DecFlat::make(C, O);
    for (int i = tid; i < 3 * EMBP; i += nthr) {
        const int d = i / EMBP, c = i % EMBP;
        sm[L::B + i] = c < EMB ? flat[f.B + d * EMB + c] : 0.0f;
    }
    stage_matrix(sm, L::W0, EMBP, HID, tid, nthr, [&](int o, int pos) { return pos < EMB ? flat[f.W[0] + o * EMB + pos] : 0.0f; });
    stage_matrix(sm, L::W3E, EMBP, HID, tid, nthr, [&](int o, int pos) { return pos < EMB ? flat[f.W[3] + o * (EMB + HID) + pos] : 0.0f; });
    stage_matrix(sm, L::W1, HID, HID, tid, nthr, [&](int o, int pos) { return flat[f.W[1] + o * HID + perm16(pos)]; });
    stage_matrix(sm, L::W2, HID, HID, tid, nthr, [&](int o, int pos) { return flat[f.W[2] + o * HID + perm16(pos)]; });
    stage_matrix(sm, L::W3H, HID, HID, tid, nthr, [&](int o, int pos) { return flat[f.W[3] + o * (EMB + HID) + EMB + perm16(pos)]; });
    stage_matrix(sm, L::W4, HID, HID, tid, nthr, [&](int o, int pos) { return flat[f.W[4] + o * HID + perm16(pos)]; });
    for (int l = 0; l < 4; ++l)      // G_l [32][C] at comp[l * 32 * C]
        stage_matrix(sm, L::FC + l * L::FCS, C, HID, tid, nthr, [&](int o, int pos) { return comp[(l * HID + o) * C + fc_channel_fwd(pos)]; });
    for (int i = tid; i < 4 * HID; i += nthr) sm[L::WO + i] = (i / HID) < O ? flat[f.Wo + i] : 0.0f;
    for (int i = tid; i < 5 * HID; i += nthr) { sm[L::BIAS + i] = comp[4 * HID * C + i]; sm[L::BIASC + i] = 0.0f; }       // b'[5][32]
    for (int i = tid; i < 4 * C; i += nthr) sm[L::WOC + i] = comp[4 * HID * C + 5 * HID + i];                            // woc[4][C]
    if (tid < 4) sm[L::BO + tid] = comp[4 * HID * C + 5 * HID + 4 * C + tid];                                            // boc[4]
}

__device__ __forceinline__ const uint32_t* wmat(const float* sm, int off) { return reinterpret_cast<const uint32_t*>(sm + off); }

// Gather the thread's 8 channels (8t..8t+7) of the trilinear feature for one sample: one 256-bit load per corner, the four
// lanes of a quad together fetch the corner's whole 128-byte line.
__device__ __forceinline__ void gather8(const GridView& G, const Bound& bnd, const float (&p)[3], int t, float* c) {
    Tri s;
    tri_setup(G, bnd, p, s);
#pragma unroll
    for (int i = 0; i < 8; ++i) c[i] = 0.0f;
    const char* const base = reinterpret_cast<const char*>(G.data + 8 * t);    // + voxel index * 128 bytes (CDIM floats per voxel)
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        unsigned idx;
        const float w = tri_corner_idx(G, s, k, idx);
        float v[8];
        ldg8(reinterpret_cast<const float*>(base + (unsigned long long)idx * (CDIM * 4)), v);
#pragma unroll
        for (int i = 0; i < 8; ++i) c[i] = fmaf(v[i], w, c[i]);
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

// Pre-activation of layer 0 (acc0) and the embedding part of the skip layer (accS), sharing the sines.  Per k-step the
// thread evaluates features 16kk + 4t .. +3 of its two samples.
template <int C, bool P3, bool STASH>
__device__ __forceinline__ void embed_layers(const float* __restrict__ sm, const float (&p)[2][3], int g, int t,
                                             float (&acc0)[4][4], float (&accS)[4][4], float* st0, float* st1) {
    using L = DecSmem<C>;
    init_bias(acc0, sm + L::BIAS + 0 * HID, t);
    init_bias(accS, sm + L::BIAS + 3 * HID, t);
#pragma unroll EMB_UNROLL
    for (int kk = 0; kk < EMBP / 16; ++kk) {
        const int f0 = 16 * kk + 4 * t;
        const float4 B0 = *reinterpret_cast<const float4*>(sm + L::B + f0);
        const float4 B1 = *reinterpret_cast<const float4*>(sm + L::B + EMBP + f0);
        const float4 B2 = *reinterpret_cast<const float4*>(sm + L::B + 2 * EMBP + f0);
        float e[2][4];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            e[r][0] = ff_sin(fmaf(p[r][2], B2.x, fmaf(p[r][1], B1.x, p[r][0] * B0.x)));
            e[r][1] = ff_sin(fmaf(p[r][2], B2.y, fmaf(p[r][1], B1.y, p[r][0] * B0.y)));
            e[r][2] = ff_sin(fmaf(p[r][2], B2.z, fmaf(p[r][1], B1.z, p[r][0] * B0.z)));
            e[r][3] = ff_sin(fmaf(p[r][2], B2.w, fmaf(p[r][1], B1.w, p[r][0] * B0.w)));
        }
        if (STASH) {
            *reinterpret_cast<float4*>(st0 + stash::E + f0) = make_float4(e[0][0], e[0][1], e[0][2], e[0][3]);
            *reinterpret_cast<float4*>(st1 + stash::E + f0) = make_float4(e[1][0], e[1][1], e[1][2], e[1][3]);
        }
        AFrag<P3> a;
        a.set(e[0][0], e[0][1], e[0][2], e[0][3], e[1][0], e[1][1], e[1][2], e[1][3]);
        kstep_fwd<P3, 4>(acc0, a, wmat(sm, L::W0), L::SE, kk, g, t);
        kstep_fwd<P3, 4>(accS, a, wmat(sm, L::W3E), L::SE, kk, g, t);
    }
}

// h += c . Fc_l^T   (c in the gather layout: c[r][4kk + i] is slot i of k-step kk); the A fragments of c are split once
// (ca) and reused by all five layers.
template <int C, bool P3>
__device__ __forceinline__ void add_cterm(const float* __restrict__ sm, int l, const AFrag<P3>* ca, int g, int t, float (&h)[4][4]) {
    using L = DecSmem<C>;
#pragma unroll
    for (int kk = 0; kk < C / 16; ++kk) kstep_fwd<P3, 4>(h, ca[kk], wmat(sm, L::FC + l * L::FCS), L::SC, kk, g, t);
}

// h = relu(acc) and the 16-bit pattern of the active units (bit 4j+q), two instructions per value: FMNMX for the relu and one
// funnel shift that collects the SIGN bit (the pattern is the complement of the sign bits; a pre-activation of exactly +0 counts
// as active, which only matters for the sub-gradient at the kink).
__device__ __forceinline__ uint32_t relu_mask(float (&h)[4][4], const float (&acc)[4][4]) {
    uint32_t m = 0;
#pragma unroll
    for (int j = 3; j >= 0; --j)
#pragma unroll
        for (int q = 3; q >= 0; --q) {
            m = __funnelshift_l(__float_as_uint(acc[j][q]), m, 1);
            h[j][q] = fmaxf(acc[j][q], 0.0f);
        }
    return ~m & 0xffffu;
}

// C-layout tile (rows g / g+8, features 8j+2t, 8j+2t+1) -> 32 floats of two stash rows
__device__ __forceinline__ void stash_tile(float* st0, float* st1, int off, const float (&h)[4][4], int t) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        __stcs(reinterpret_cast<float2*>(st0 + off + 8 * j + 2 * t), make_float2(h[j][0], h[j][1]));   // streaming: the stash is written once, read once
        __stcs(reinterpret_cast<float2*>(st1 + off + 8 * j + 2 * t), make_float2(h[j][2], h[j][3]));
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
    AFrag<P3> ca[C / 16];
#pragma unroll
    for (int kk = 0; kk < C / 16; ++kk)
        ca[kk].set(c[0][4 * kk], c[0][4 * kk + 1], c[0][4 * kk + 2], c[0][4 * kk + 3], c[1][4 * kk], c[1][4 * kk + 1], c[1][4 * kk + 2], c[1][4 * kk + 3]);
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
            for (int kk = 0; kk < 2; ++kk) {
                AFrag<P3> a;
                afrag_from_c<P3>(a, h[2 * kk], h[2 * kk + 1]);
                kstep_fwd<P3, 4>(acc, a, wmat(sm, L::w(i)), L::SH, kk, g, t);
            }
        }
        masks[i] = relu_mask(h, acc);
        add_bias(h, sm + L::BIASC + i * HID, t);
        add_cterm<C, P3>(sm, i, ca, g, t, h);
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

// Same on the composed image (stage_decoder_composed): the grid-feature term of every layer is accumulated into that layer's
// pre-activation BEFORE the previous layer's relu is needed, so those MMAs run while the relu / fp16 split of the previous layer
// occupies the ALUs.  No stash variant: the weight gradient needs the uncomposed block outputs.
template <int C, int O, bool P3>
__device__ __forceinline__ void decoder_forward_composed(const float* __restrict__ sm, const float (&p)[2][3],
                                                         const float (&c)[2][C / 4], int g, int t, float (&out)[2][4],
                                                         uint32_t (&masks)[5]) {
    using L = DecSmem<C>;
    float acc[4][4], accS[4][4], nxt[4][4], h[4][4];
    AFrag<P3> ca[C / 16];
#pragma unroll
    for (int kk = 0; kk < C / 16; ++kk)
        ca[kk].set(c[0][4 * kk], c[0][4 * kk + 1], c[0][4 * kk + 2], c[0][4 * kk + 3], c[1][4 * kk], c[1][4 * kk + 1], c[1][4 * kk + 2], c[1][4 * kk + 3]);
    // grid-feature part of the output layer, per lane over its channels (finished by the quad sum below)
    constexpr int NO = O == 4 ? 3 : 1;
    float oc[2][NO];
#pragma unroll
    for (int o = 0; o < NO; ++o) {
        oc[0][o] = oc[1][o] = 0.0f;
#pragma unroll
        for (int b = 0; b < C / 32; ++b)
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float w = sm[L::WOC + o * C + 32 * b + 8 * t + i];
                oc[0][o] = fmaf(c[0][8 * b + i], w, oc[0][o]); oc[1][o] = fmaf(c[1][8 * b + i], w, oc[1][o]);
            }
    }
    embed_layers<C, P3, false>(sm, p, g, t, acc, accS, nullptr, nullptr);      // acc = b'_0 + W0 e, accS = b'_3 + W3e e
    add_cterm<C, P3>(sm, 2, ca, g, t, accS);                                    // + G_2 c
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        // pre-activation of the NEXT layer starts with its bias and grid-feature term: independent of this layer's relu
        if (i == 0 || i == 1 || i == 3) {
            init_bias(nxt, sm + L::BIAS + (i + 1) * HID, t);
            add_cterm<C, P3>(sm, i, ca, g, t, nxt);                             // G_i c
        }
        masks[i] = relu_mask(h, acc);
        if (i < 4) {
            float (&dst)[4][4] = (i == 2) ? accS : nxt;
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) {
                AFrag<P3> a;
                afrag_from_c<P3>(a, h[2 * kk], h[2 * kk + 1]);
                kstep_fwd<P3, 4>(dst, a, wmat(sm, L::w(i + 1)), L::SH, kk, g, t);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int q = 0; q < 4; ++q) acc[j][q] = dst[j][q];
        }
    }
#pragma unroll
    for (int o = 0; o < NO; ++o) {
        float s0 = oc[0][o], s1 = oc[1][o];
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
    static constexpr int SH = wstride(HID);
    static constexpr int W0 = 0;                 // [32][SH], positions in the gather order
    static constexpr int W1 = W0 + HID * SH, W2 = W1 + HID * SH;
    static constexpr int W3C = W2 + HID * SH;    // skip, c columns (gather order)
    static constexpr int W3H = W3C + HID * SH, W4 = W3H + HID * SH;
    static constexpr int WO = W4 + HID * SH;     // [32] fp32
    static constexpr int BIAS = WO + 32;         // b[5][32]
    static constexpr int BO = BIAS + 160;
    static constexpr int TOTAL = BO + 4;
};
__device__ inline void stage_coarse(float* sm, const float* __restrict__ flat, int tid, int nthr) {
    using L = CoarseSmem;
    const CoarseFlat f = CoarseFlat::make();
    stage_matrix(sm, L::W0, HID, HID, tid, nthr, [&](int o, int pos) { return flat[f.W[0] + o * CDIM + fc_channel_fwd(pos)]; });
    stage_matrix(sm, L::W1, HID, HID, tid, nthr, [&](int o, int pos) { return flat[f.W[1] + o * HID + perm16(pos)]; });
    stage_matrix(sm, L::W2, HID, HID, tid, nthr, [&](int o, int pos) { return flat[f.W[2] + o * HID + perm16(pos)]; });
    stage_matrix(sm, L::W3C, HID, HID, tid, nthr, [&](int o, int pos) { return flat[f.W[3] + o * (CDIM + HID) + fc_channel_fwd(pos)]; });
    stage_matrix(sm, L::W3H, HID, HID, tid, nthr, [&](int o, int pos) { return flat[f.W[3] + o * (CDIM + HID) + CDIM + perm16(pos)]; });
    stage_matrix(sm, L::W4, HID, HID, tid, nthr, [&](int o, int pos) { return flat[f.W[4] + o * HID + perm16(pos)]; });
    for (int i = tid; i < HID; i += nthr) sm[L::WO + i] = flat[f.Wo + i];
    for (int i = tid; i < 5 * HID; i += nthr) sm[L::BIAS + i] = flat[f.b[i / HID] + i % HID];
    if (tid == 0) sm[L::BO] = flat[f.bo];
}
template <bool P3>
__device__ __forceinline__ void coarse_forward(const float* __restrict__ sm, const float (&c)[2][8], int g, int t, float (&out)[2], uint32_t (&masks)[5]) {
    using L = CoarseSmem;
    float acc[4][4], h[4][4];
    const int wofs[5] = {L::W0, L::W1, L::W2, L::W3H, L::W4};
    AFrag<P3> ca[2];
#pragma unroll
    for (int kk = 0; kk < 2; ++kk)
        ca[kk].set(c[0][4 * kk], c[0][4 * kk + 1], c[0][4 * kk + 2], c[0][4 * kk + 3], c[1][4 * kk], c[1][4 * kk + 1], c[1][4 * kk + 2], c[1][4 * kk + 3]);
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        init_bias(acc, sm + L::BIAS + i * HID, t);
        if (i == 0 || i == 3) {
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) kstep_fwd<P3, 4>(acc, ca[kk], wmat(sm, i == 0 ? L::W0 : L::W3C), L::SH, kk, g, t);
        }
        if (i > 0) {
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) {
                AFrag<P3> a;
                afrag_from_c<P3>(a, h[2 * kk], h[2 * kk + 1]);
                kstep_fwd<P3, 4>(acc, a, wmat(sm, wofs[i]), L::SH, kk, g, t);
            }
        }
        masks[i] = relu_mask(h, acc);
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

// Backward image of the coarse decoder: the transposed matrices (rows = input feature, contraction over the output features,
// whose operand comes from accumulator fragments -> perm16); the two matrices that multiply the grid feature (W0, the c columns
// of the skip layer) have their rows in the scatter order of fc_channel_bwd.  Same slots as CoarseSmem.
__device__ inline void stage_coarse_bwd(float* sm, const float* __restrict__ flat, int tid, int nthr) {
    using L = CoarseSmem;
    const CoarseFlat f = CoarseFlat::make();
    stage_matrix(sm, L::W0, HID, HID, tid, nthr, [&](int n, int pos) { return flat[f.W[0] + perm16(pos) * CDIM + fc_channel_bwd(n)]; });
    stage_matrix(sm, L::W1, HID, HID, tid, nthr, [&](int n, int pos) { return flat[f.W[1] + perm16(pos) * HID + n]; });
    stage_matrix(sm, L::W2, HID, HID, tid, nthr, [&](int n, int pos) { return flat[f.W[2] + perm16(pos) * HID + n]; });
    stage_matrix(sm, L::W3C, HID, HID, tid, nthr, [&](int n, int pos) { return flat[f.W[3] + perm16(pos) * (CDIM + HID) + fc_channel_bwd(n)]; });
    stage_matrix(sm, L::W3H, HID, HID, tid, nthr, [&](int n, int pos) { return flat[f.W[3] + perm16(pos) * (CDIM + HID) + CDIM + n]; });
    stage_matrix(sm, L::W4, HID, HID, tid, nthr, [&](int n, int pos) { return flat[f.W[4] + perm16(pos) * HID + n]; });
    for (int i = tid; i < HID; i += nthr) sm[L::WO + i] = flat[f.Wo + i];
}

// Data gradient of MLP_no_xyz (MLP.cpp:165-181) for one tile: gout[r] = cotangent of the occupancy of rows g / g+8, masks = relu
// patterns of the training forward; returns gcf[r][8] = d L / d c for the thread's channels 4t..4t+3, 16+4t..16+4t+3.
template <bool P3>
__device__ __forceinline__ void coarse_backward(const float* __restrict__ sm, int g, int t, const float (&gout)[2], const uint32_t (&masks)[5],
                                                float (&gcf)[2][8]) {
    using L = CoarseSmem;
    float gh[4][4], gu[4][4], gc[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float2 w = *reinterpret_cast<const float2*>(sm + L::WO + 8 * j + 2 * t);
        gh[j][0] = gout[0] * w.x; gh[j][1] = gout[0] * w.y; gh[j][2] = gout[1] * w.x; gh[j][3] = gout[1] * w.y;
        gc[j][0] = gc[j][1] = gc[j][2] = gc[j][3] = 0.0f;
    }
    const int wofs[5] = {L::W0, L::W1, L::W2, L::W3H, L::W4};
#pragma unroll
    for (int i = 4; i >= 0; --i) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int q = 0; q < 4; ++q) gu[j][q] = ((masks[i] >> (4 * j + q)) & 1u) ? gh[j][q] : 0.0f;
        if (i > 0) {
#pragma unroll
            for (int j = 0; j < 4; ++j) gh[j][0] = gh[j][1] = gh[j][2] = gh[j][3] = 0.0f;
        }
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {
            AFrag<P3> a;
            afrag_from_c<P3>(a, gu[2 * kk], gu[2 * kk + 1]);
            if (i > 0) kstep_fwd<P3, 4>(gh, a, wmat(sm, wofs[i]), L::SH, kk, g, t);          // g_h = g_u W_i (hidden columns)
            if (i == 3) kstep_fwd<P3, 4>(gc, a, wmat(sm, L::W3C), L::SH, kk, g, t);          // skip layer: the c columns
            if (i == 0) kstep_fwd<P3, 4>(gc, a, wmat(sm, L::W0), L::SH, kk, g, t);           // first layer acts on c itself
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        gcf[0][2 * j] = gc[j][0]; gcf[0][2 * j + 1] = gc[j][1];
        gcf[1][2 * j] = gc[j][2]; gcf[1][2 * j + 1] = gc[j][3];
    }
}

}  // namespace nsb
