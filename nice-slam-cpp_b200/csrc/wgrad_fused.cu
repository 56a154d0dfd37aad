// Colour-decoder weight gradient WITHOUT an HBM stash (replaces the round-1 pair "forward/backward stash 712 floats per sample +
// split-K k_wgrad": 1.2 GB of HBM traffic per colour iteration for a 15.9 k-float result).
//
// Every parameter gradient of the decoder is a sum over samples of an outer product  dW_i = sum_s gu_i[s]^T x_i[s]  (gu_i = gradient
// at the relu output of layer i, x_i = input of layer i).  The kernel RECOMPUTES both sides per 16-sample tile from what the
// iteration already has in HBM (rays, z values, the raw cotangent, the relu masks: ~100 bytes per sample) and contracts them on the
// tensor cores with the SAMPLE index as the k dimension, handing the operands between warps through shared memory:
//
//   round = 16 tiles, one per warp.   Every warp is a producer (recomputes its tile) and a consumer (owns some dW blocks).
//   P1  gu chain from the saved masks:  g_h5 = g_out Wo, gu_i = mask_i(g_h), g_h = gu_i W_i  -> all gu_i of the tile are parked in
//       shared memory as fp16 rows [sample][feature] (G tile; power-of-two scaled, single fp16: the gradient tolerance is 1e-3)
//   P2  forward (uncomposed, fp32-grade fp16 hi/lo split as everywhere else), layer by layer; after each layer the tile's x_i
//       (e chunks, h_1..h_5, c, and q = g_e cos(pB) for dB) goes to a one-chunk staging buffer (X tile, fp16 hi + lo planes)
//   C   all warps: D[m-tile of G features][n-tile of X features] += sum over the round's 16 tiles of  G^T X  -- both fragments
//       come straight out of the row-major tiles with ldmatrix.trans (k = sample index), 2 MMAs per product (G x X_hi, G x X_lo);
//       the round's partial sums are added to the gradient arena with vector reductions (a CTA runs ~5 rounds).
//   The weight matrices live ONCE in shared memory as plain row-major fp16 hi / lo planes (row stride 2K + 16 bytes): the forward
//   reads its B fragments with ldmatrix, the backward chain reads the TRANSPOSED fragments of the same bytes with ldmatrix.trans.
//   The grid-feature weights obey  dFc_i = W_{i+1}^T (sum_s gu_{i+1}^T c),  dbc_i = W_{i+1}^T db_{i+1}  (g_h{i+1} = gu_{i+1} W_{i+1}),
//   so only the 32 x 32 products M_i = sum gu_{i+1}^T c are accumulated here; k_wgrad_finish applies W^T once per iteration.
// This is the autograd weight gradient of MLP::forward (MLP.cpp:76-102) that loss.backward() produces at Mapper.cpp:444.
#include "decode.cuh"
#include "params.h"

namespace nsb {
namespace wgf {

constexpr int NW = 16;                        // warps per CTA
constexpr int GW = 8;                         // warps per GROUP = tiles per round: the CTA runs two independent groups (own staging
                                              // buffers, own named barriers) that drift out of phase, so that one group's barrier
                                              // waits are filled with the other group's work
constexpr int THREADS = NW * 32;
constexpr int RS32 = 80, RS96 = 208;          // row strides (bytes) of the K = 32 / K = 96 planes: 2K + 16 -> conflict-free ldmatrix
constexpr int PL32 = 32 * RS32, PL96 = 32 * RS96;
// weight image (bytes): hi plane followed by lo plane for every matrix
constexpr int I_W0 = 0;
constexpr int I_W3E = I_W0 + 2 * PL96;
constexpr int I_WH = I_W3E + 2 * PL96;        // W1, W2, W3H, W4
constexpr int I_FC = I_WH + 4 * 2 * PL32;     // Fc_0 .. Fc_4
constexpr int I_F32 = I_FC + 5 * 2 * PL32;    // fp32 section
constexpr int F_B = 0, F_b = 3 * EMBP, F_bc = F_b + 5 * HID, F_Wo = F_bc + 5 * HID, F_bo = F_Wo + 4 * HID, F_N = F_bo + 4;
constexpr int IMG_BYTES = (I_F32 + 4 * F_N + 15) & ~15;
// G tile: one row per sample, fp16: gu_0..gu_4 at columns 32 i, g_out at 160..162, p_hi at 168..170, p_lo at 171..173
constexpr int GCOLS = 184, GS = 2 * GCOLS;    // 368 bytes per row = 112 (mod 128)
constexpr int G_TILE = 16 * GS;
constexpr int C_GOUT = 160, C_P = 168;
// X tile: one 32-wide chunk, hi plane then lo plane, rows of 80 bytes
constexpr int X_PLANE = 16 * RS32, X_TILE = 2 * X_PLANE;
constexpr int S_IMG = 0, S_G = IMG_BYTES, S_X = S_G + NW * G_TILE, S_FLAG = S_X + NW * X_TILE, SMEM_BYTES = S_FLAG + 64;
constexpr float GSCALE = 64.0f, GINV = 1.0f / 64.0f;   // exact power of two: moves the gradients away from the fp16 subnormals

// scratch accumulators outside the gradient arena (floats): M_0..M_3 [32][32] (rows = outputs of layer i+1, cols = channel), Mo [4][32]
constexpr int SCR_M = 0, SCR_MO = 4 * HID * HID, SCR_N = SCR_MO + 4 * HID;

// image column of the K = 96 planes that holds embedding feature f, and of the Fc planes that holds channel ch: chosen so that the
// producing thread's own values are the A-fragment slots (see embed / cterm below)
__host__ __device__ __forceinline__ int emb_col(int f) { const int kk = f >> 4, r = f & 15, t = r >> 2, h = (r >> 1) & 1, i = r & 1; return 16 * kk + 8 * h + 2 * t + i; }
__host__ __device__ __forceinline__ int ch_col(int ch) { const int b = ch >> 5, c = ch & 31, t = c >> 3, kk = (c >> 2) & 1, h = (c >> 1) & 1, i = c & 1; return 32 * b + 16 * kk + 8 * h + 2 * t + i; }

__device__ __forceinline__ void put_hl(uint8_t* img, int mat, int plane, int rs, int row, int col, float v) {
    const __half h = __float2half_rn(v), l = __float2half_rn(v - __half2float(h));
    *reinterpret_cast<__half*>(img + mat + row * rs + 2 * col) = h;
    *reinterpret_cast<__half*>(img + mat + plane + row * rs + 2 * col) = l;
}

// Builds the image in global memory (once per update of the colour decoder); the kernel copies it with 16-byte loads.
__global__ void __launch_bounds__(512) k_build_wgimg(const float* __restrict__ flat, uint8_t* __restrict__ img) {
    const DecFlat f = DecFlat::make(32, 4);
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nthr = gridDim.x * blockDim.x;
    for (int i = tid; i < HID * EMBP; i += nthr) {
        const int o = i / EMBP, ft = i % EMBP;
        put_hl(img, I_W0, PL96, RS96, o, emb_col(ft), ft < EMB ? flat[f.W[0] + o * EMB + ft] : 0.0f);
        put_hl(img, I_W3E, PL96, RS96, o, emb_col(ft), ft < EMB ? flat[f.W[3] + o * (EMB + HID) + ft] : 0.0f);
    }
    for (int i = tid; i < 4 * HID * HID; i += nthr) {
        const int l = i / (HID * HID), o = (i / HID) % HID, k = i % HID;    // l = 0..3 <-> W1, W2, W3H, W4
        const float v = l == 2 ? flat[f.W[3] + o * (EMB + HID) + EMB + k] : flat[f.W[l + 1] + o * HID + k];
        put_hl(img, I_WH + l * 2 * PL32, PL32, RS32, o, k, v);
    }
    for (int i = tid; i < 5 * HID * HID; i += nthr) {
        const int l = i / (HID * HID), o = (i / HID) % HID, ch = i % HID;
        put_hl(img, I_FC + l * 2 * PL32, PL32, RS32, o, ch_col(ch), flat[f.Fc[l] + o * 32 + ch]);
    }
    float* fs = reinterpret_cast<float*>(img + I_F32);
    for (int i = tid; i < 3 * EMBP; i += nthr) { const int d = i / EMBP, c = i % EMBP; fs[F_B + i] = c < EMB ? flat[f.B + d * EMB + c] : 0.0f; }
    for (int i = tid; i < 5 * HID; i += nthr) { fs[F_b + i] = flat[f.b[i / HID] + i % HID]; fs[F_bc + i] = flat[f.bc[i / HID] + i % HID]; }
    for (int i = tid; i < 4 * HID; i += nthr) fs[F_Wo + i] = flat[f.Wo + i];
    if (tid < 4) fs[F_bo + tid] = flat[f.bo + tid];
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ldsm4(uint32_t addr, uint32_t (&r)[4]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm4t(uint32_t addr, uint32_t (&r)[4]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void bar_sync(int id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(GW * 32) : "memory"); }

// Per-lane byte offsets of the ldmatrix row addresses, computed once per thread: forward (rows of the plane = outputs, k contiguous)
// and transposed (contraction over the ROWS of the same plane) orientation, for the K = 32 and K = 96 planes.
struct LaneOff { uint32_t f32, f96, t32, t96; };
__device__ __forceinline__ LaneOff lane_offsets(int lane) {
    LaneOff o;
    const uint32_t hl = (lane >> 4) & 1, kh = (lane >> 3) & 1, r = lane & 7;
    o.f32 = hl * PL32 + r * RS32 + kh * 16; o.f96 = hl * PL96 + r * RS96 + kh * 16;
    o.t32 = hl * PL32 + (kh * 8 + r) * RS32; o.t96 = hl * PL96 + (kh * 8 + r) * RS96;
    return o;
}
// acc[j] += A(kk) . W[8j..8j+7][k-step kk]^T, fp32-grade; `base` = matrix + the thread's forward lane offset
__device__ __forceinline__ void kstep_w(float (&acc)[4][4], const AFrag<true>& a, uint32_t base, int rs, int kk) {
    base += 32 * kk;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        uint32_t b[4];
        ldsm4(base + 8 * j * rs, b);
        mma_f16(acc[j], a.lo, b[0], b[1]);
        mma_f16(acc[j], a.hi, b[2], b[3]);
        mma_f16(acc[j], a.hi, b[0], b[1]);
    }
}
// acc[j] += A(kk) . W[16kk..16kk+15][8(j0+j)..]  (transposed load); `base` = matrix + the thread's transposed lane offset
template <bool P3>
__device__ __forceinline__ void kstep_wt(float (&acc)[4][4], const AFrag<P3>& a, uint32_t base, int rs, int kk, int j0) {
    base += 16 * kk * rs + 16 * j0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        uint32_t b[4];
        ldsm4t(base + 16 * j, b);
        if (P3) mma_f16(acc[j], a.lo, b[0], b[1]);
        mma_f16(acc[j], a.hi, b[2], b[3]);
        mma_f16(acc[j], a.hi, b[0], b[1]);
    }
}

// C-layout tile (rows g / g+8, features 8j+2t, 8j+2t+1) -> X tile chunk (hi + lo planes) at natural columns
__device__ __forceinline__ void stage_x(uint8_t* xt, const float (&h)[4][4], int g, int t) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        uint32_t hi, lo;
        split_f16(h[j][0], h[j][1], hi, lo);
        *reinterpret_cast<uint32_t*>(xt + g * RS32 + 2 * (8 * j + 2 * t)) = hi;
        *reinterpret_cast<uint32_t*>(xt + X_PLANE + g * RS32 + 2 * (8 * j + 2 * t)) = lo;
        split_f16(h[j][2], h[j][3], hi, lo);
        *reinterpret_cast<uint32_t*>(xt + (g + 8) * RS32 + 2 * (8 * j + 2 * t)) = hi;
        *reinterpret_cast<uint32_t*>(xt + X_PLANE + (g + 8) * RS32 + 2 * (8 * j + 2 * t)) = lo;
    }
}
// masked gradient tile -> G tile columns col0.. (single fp16, scaled)
__device__ __forceinline__ void stage_g(uint8_t* gt, int col0, const float (&gu)[4][4], int g, int t) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        *reinterpret_cast<uint32_t*>(gt + g * GS + 2 * (col0 + 8 * j + 2 * t)) = pack_f16(gu[j][0] * GSCALE, gu[j][1] * GSCALE);
        *reinterpret_cast<uint32_t*>(gt + (g + 8) * GS + 2 * (col0 + 8 * j + 2 * t)) = pack_f16(gu[j][2] * GSCALE, gu[j][3] * GSCALE);
    }
}

// One (m-tile, n-tile) block of the round: d = sum over the active tiles of  G[:, mcol..mcol+15]^T . X[:, 8 nt..8 nt+7]
// ga / xb: G / X staging base + the thread's ldmatrix lane offset (consume_offsets); tiles [t0, t0 + NT) of the round, `act` = bit mask
// of the live tiles.  Two independent accumulation chains (X_hi and X_lo products) keep the tensor pipe busy.
template <int NT>
__device__ __forceinline__ void consume(uint32_t ga, uint32_t xb, uint32_t act, int t0, int mcol, int nt, float (&d)[4]) {
    float e[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    d[0] = d[1] = d[2] = d[3] = 0.0f;
    ga += t0 * G_TILE + 2 * mcol; xb += t0 * X_TILE + 16 * nt;
#pragma unroll
    for (int tile = 0; tile < NT; ++tile) {
        if ((act >> (t0 + tile)) & 1u) {
            uint32_t a[4], b[4];
            ldsm4t(ga + tile * G_TILE, a);
            ldsm4t(xb + tile * X_TILE, b);
            mma_f16(d, a, b[0], b[1]);
            mma_f16(e, a, b[2], b[3]);
        }
    }
    d[0] += e[0]; d[1] += e[1]; d[2] += e[2]; d[3] += e[3];
}
// column sums of an m-tile of G over the round (bias gradients): B = ones
__device__ __forceinline__ void consume_ones(uint32_t ga, uint32_t act, int mcol, float (&d)[4]) {
    d[0] = d[1] = d[2] = d[3] = 0.0f;
    ga += 2 * mcol;
#pragma unroll
    for (int tile = 0; tile < GW; ++tile) {
        if ((act >> tile) & 1u) {
            uint32_t a[4];
            ldsm4t(ga + tile * G_TILE, a);
            mma_f16(d, a, 0x3C003C00u, 0x3C003C00u);
        }
    }
}
__device__ __forceinline__ void red2(float* p, float a, float b) { asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(a), "f"(b) : "memory"); }
// adds the block (rows = G features mrow0 + g, + 8; columns 8 nt + 2t, +1) into dst[row * ld + col], columns < ncol only
__device__ __forceinline__ void flush_block(float* dst, int ld, int mrow0, int nrows, int col0, int ncol, const float (&d)[4], float scale, int g, int t) {
    const int c = col0 + 2 * t;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int row = mrow0 + g + 8 * h;
        if (row >= nrows) continue;
        float* p = dst + (size_t)row * ld + c;
        if (c + 1 < ncol && (((size_t)p & 7) == 0)) red2(p, d[2 * h] * scale, d[2 * h + 1] * scale);
        else {
            if (c < ncol) atomicAdd(p, d[2 * h] * scale);
            if (c + 1 < ncol) atomicAdd(p + 1, d[2 * h + 1] * scale);
        }
    }
}

struct Params {
    DecodeParams D;             // rays, z, valid, g_raw, masks, colour grid, bound
    const uint8_t* img;         // k_build_wgimg output
    float* dflat;               // gradient of the colour decoder's flat parameter vector (gradient arena)
    float* scratch;             // [SCR_N] M_i / Mo accumulators (zero on entry, consumed and cleared by k_wgrad_finish)
};

__global__ void __launch_bounds__(THREADS, 1) k_wgrad_fused(const Params Q) {
    extern __shared__ __align__(128) uint8_t sm[];
    const DecodeParams& P = Q.D;
    const int tid = threadIdx.x;
    int warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    // the thread-invariant indices and ldmatrix offsets below are made opaque to the compiler: at 128 registers it would otherwise
    // REMATERIALISE them from threadIdx at every use (measured: ~45 M of 133 M executed instructions were such recomputations)
    asm volatile("" : "+r"(warp), "+r"(lane), "+r"(g), "+r"(t));
    {
        const uint4* src = reinterpret_cast<const uint4*>(Q.img);
        uint4* dst = reinterpret_cast<uint4*>(sm + S_IMG);
        for (int i = tid; i < IMG_BYTES / 16; i += THREADS) dst[i] = __ldg(src + i);
    }
    __syncthreads();
    const uint32_t sb = smem_u32(sm);
    const float* fs = reinterpret_cast<const float*>(sm + S_IMG + I_F32);
    int grp = warp >> 3, wl = warp & 7;           // group of the warp, warp within the group
    asm volatile("" : "+r"(grp), "+r"(wl));
    volatile int* active = reinterpret_cast<volatile int*>(sm + S_FLAG) + GW * grp;
    uint8_t* gt = sm + S_G + warp * G_TILE;
    uint8_t* xt = sm + S_X + warp * X_TILE;
    const DecFlat f = DecFlat::make(32, 4);
    const int ntiles = P.P / TILE;
    const int nrounds = (ntiles + GW - 1) / GW;
    const int bar1 = 1 + 2 * grp, bar2 = 2 + 2 * grp;
    LaneOff lo = lane_offsets(lane);
    const uint32_t img = sb + S_IMG;
    lo.f32 += img; lo.f96 += img; lo.t32 += img; lo.t96 += img;
    // ldmatrix lane offsets of the consumer side: G rows (samples) x feature columns, X planes
    uint32_t ga = sb + S_G + grp * GW * G_TILE + (((lane >> 4) & 1) * 8 + (lane & 7)) * GS + ((lane >> 3) & 1) * 16;
    uint32_t xb = sb + S_X + grp * GW * X_TILE + ((lane >> 4) & 1) * X_PLANE + (((lane >> 3) & 1) * 8 + (lane & 7)) * RS32;
    // own G tile, non-transposed A fragments (Q phases): rows (lane>>3 & 1)*8 + r, columns (lane>>4)*8
    uint32_t gown = sb + S_G + warp * G_TILE + (((lane >> 3) & 1) * 8 + (lane & 7)) * GS + ((lane >> 4) & 1) * 16;
    asm volatile("" : "+r"(lo.f32), "+r"(lo.f96), "+r"(lo.t32), "+r"(lo.t96), "+r"(ga), "+r"(xb), "+r"(gown));

    for (int round = 2 * blockIdx.x + grp; round < nrounds; round += 2 * gridDim.x) {
        const int tile = round * GW + wl;
        // ------------------------------------------------------------------ inputs of the tile
        bool live = tile < ntiles;
        const int base = tile * TILE, s0 = base + 2 * g, s1 = s0 + 1;
        float p[2][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};
        AFrag<true> ca[2];
        float acc[4][4], accS[4][4], h[4][4];
        if (live) {
            const int ray = base / P.S;
            const uint8_t ok = P.valid ? P.valid[ray] : (uint8_t)1;
            const float o[3] = {P.rays_o[3 * ray], P.rays_o[3 * ray + 1], P.rays_o[3 * ray + 2]};
            const float d[3] = {P.rays_d[3 * ray], P.rays_d[3 * ray + 1], P.rays_d[3 * ray + 2]};
            const float zz[2] = {P.z[s0], P.z[s1]};
            const float4 gr0 = *reinterpret_cast<const float4*>(P.g_raw + 4 * (size_t)s0), gr1 = *reinterpret_cast<const float4*>(P.g_raw + 4 * (size_t)s1);
            const uint32_t* mb = P.masks + ((size_t)2 * ntiles + tile) * 96 + lane;      // decoder 3's slot of the packed relu masks
            const uint32_t m0 = mb[0], m1 = mb[32], m2 = mb[64];
            const bool any = gr0.x != 0.f || gr0.y != 0.f || gr0.z != 0.f || gr1.x != 0.f || gr1.y != 0.f || gr1.z != 0.f;
            live = ok && __any_sync(0xffffffffu, any);          // nothing flows into this tile: it contributes exact zeros
            if (live) {
                const uint32_t masks[5] = {m0 & 0xffffu, m0 >> 16, m1 & 0xffffu, m1 >> 16, m2};
                const float gout[2][3] = {{gr0.x, gr0.y, gr0.z}, {gr1.x, gr1.y, gr1.z}};
#pragma unroll
                for (int r = 0; r < 2; ++r)
#pragma unroll
                    for (int a = 0; a < 3; ++a) p[r][a] = __fadd_rn(o[a], __fmul_rn(d[a], zz[r]));
                // ------------------------------------------------------------ P1: the gu chain, parked in the G tile
                float gh[4][4], gu[4][4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    gh[j][0] = gh[j][1] = gh[j][2] = gh[j][3] = 0.0f;
#pragma unroll
                    for (int o3 = 0; o3 < 3; ++o3) {
                        const float2 w = *reinterpret_cast<const float2*>(fs + F_Wo + o3 * HID + 8 * j + 2 * t);
                        gh[j][0] = fmaf(gout[0][o3], w.x, gh[j][0]); gh[j][1] = fmaf(gout[0][o3], w.y, gh[j][1]);
                        gh[j][2] = fmaf(gout[1][o3], w.x, gh[j][2]); gh[j][3] = fmaf(gout[1][o3], w.y, gh[j][3]);
                    }
                }
#pragma unroll
                for (int i = 4; i >= 0; --i) {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
#pragma unroll
                        for (int q = 0; q < 4; ++q) gu[j][q] = ((masks[i] >> (4 * j + q)) & 1u) ? gh[j][q] : 0.0f;
                    stage_g(gt, 32 * i, gu, g, t);
                    if (i > 0) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) gh[j][0] = gh[j][1] = gh[j][2] = gh[j][3] = 0.0f;
#pragma unroll
                        for (int kk = 0; kk < 2; ++kk) {
                            AFrag<true> a;
                            afrag_from_c<true>(a, gu[2 * kk], gu[2 * kk + 1]);
                            kstep_wt<true>(gh, a, lo.t32 + I_WH + (i - 1) * 2 * PL32, RS32, kk, 0);       // g_h = gu_i W_i
                        }
                    }
                }
                if (t == 0) {   // g_out and the position (hi, lo) rows of the "GP" m-tile
#pragma unroll
                    for (int r = 0; r < 2; ++r) {
                        __half* row = reinterpret_cast<__half*>(gt + (g + 8 * r) * GS);
#pragma unroll
                        for (int o3 = 0; o3 < 3; ++o3) {
                            row[C_GOUT + o3] = __float2half_rn(gout[r][o3] * GSCALE);
                            const __half ph = __float2half_rn(p[r][o3]);
                            row[C_P + o3] = ph; row[C_P + 3 + o3] = __float2half_rn(p[r][o3] - __half2float(ph));
                        }
#pragma unroll
                        for (int o3 = 3; o3 < 8; ++o3) row[C_GOUT + o3] = __float2half_rn(0.0f);
                        row[C_P + 6] = row[C_P + 7] = __float2half_rn(0.0f);
#pragma unroll
                        for (int o3 = 8; o3 < 16; ++o3) row[C_P + o3] = __float2half_rn(0.0f);
                    }
                }
                // ------------------------------------------------------------ grid feature of the two samples (gather layout: channels 8t..8t+7)
                float c[2][8];
                gather8(P.grid[3], P.bnd, p[0], t, c[0]);
                gather8(P.grid[3], P.bnd, p[1], t, c[1]);
#pragma unroll
                for (int kk = 0; kk < 2; ++kk)
                    ca[kk].set(c[0][4 * kk], c[0][4 * kk + 1], c[0][4 * kk + 2], c[0][4 * kk + 3], c[1][4 * kk], c[1][4 * kk + 1], c[1][4 * kk + 2], c[1][4 * kk + 3]);
                init_bias(acc, fs + F_b + 0 * HID, t);
                init_bias(accS, fs + F_b + 3 * HID, t);
            }
        }
        if (lane == 0) active[wl] = live ? 1 : 0;
        __syncwarp();
        uint32_t act = 0;

        // ==================================================================== E phases: embedding chunks, dW0 / dW3 (embedding columns)
#pragma unroll 1
        for (int k = 0; k < 3; ++k) {
            if (live) {
#pragma unroll
                for (int kq = 0; kq < 2; ++kq) {
                    const int kk = 2 * k + kq, f0 = 16 * kk + 4 * t;
                    const float4 B0 = *reinterpret_cast<const float4*>(fs + F_B + f0);
                    const float4 B1 = *reinterpret_cast<const float4*>(fs + F_B + EMBP + f0);
                    const float4 B2 = *reinterpret_cast<const float4*>(fs + F_B + 2 * EMBP + f0);
                    float e[2][4];
#pragma unroll
                    for (int r = 0; r < 2; ++r) {
                        e[r][0] = ff_sin(fmaf(p[r][2], B2.x, fmaf(p[r][1], B1.x, p[r][0] * B0.x)));
                        e[r][1] = ff_sin(fmaf(p[r][2], B2.y, fmaf(p[r][1], B1.y, p[r][0] * B0.y)));
                        e[r][2] = ff_sin(fmaf(p[r][2], B2.z, fmaf(p[r][1], B1.z, p[r][0] * B0.z)));
                        e[r][3] = ff_sin(fmaf(p[r][2], B2.w, fmaf(p[r][1], B1.w, p[r][0] * B0.w)));
                    }
                    AFrag<true> a;
                    a.set(e[0][0], e[0][1], e[0][2], e[0][3], e[1][0], e[1][1], e[1][2], e[1][3]);
                    // stage the chunk at natural feature columns (16 kq + 4t .. +3): the hi / lo words of the A fragment are exactly the pairs
                    *reinterpret_cast<uint2*>(xt + g * RS32 + 2 * (16 * kq + 4 * t)) = make_uint2(a.hi[0], a.hi[2]);
                    *reinterpret_cast<uint2*>(xt + (g + 8) * RS32 + 2 * (16 * kq + 4 * t)) = make_uint2(a.hi[1], a.hi[3]);
                    *reinterpret_cast<uint2*>(xt + X_PLANE + g * RS32 + 2 * (16 * kq + 4 * t)) = make_uint2(a.lo[0], a.lo[2]);
                    *reinterpret_cast<uint2*>(xt + X_PLANE + (g + 8) * RS32 + 2 * (16 * kq + 4 * t)) = make_uint2(a.lo[1], a.lo[3]);
                    kstep_w(acc, a, lo.f96 + I_W0, RS96, kk);
                    kstep_w(accS, a, lo.f96 + I_W3E, RS96, kk);
                }
            }
            bar_sync(bar1);
            if (k == 0) act = __ballot_sync(0xffffffffu, active[lane & 7] != 0) & 0xffu;
#pragma unroll 1
            for (int blk = wl; blk < 16; blk += GW) {   // 16 blocks: m-tiles {gu0 rows 0-15, gu0 rows 16-31, gu3 ...} x 4 n-tiles
                const int mi = blk >> 2, nt = blk & 3;
                const int mcol = (mi < 2 ? 0 : 96) + 16 * (mi & 1);
                float d[4];
                consume<GW>(ga, xb, act, 0, mcol, nt, d);
                const int ncol = EMB - 32 * k;                     // features 93..95 are padding
                if (mi < 2) flush_block(Q.dflat + f.W[0] + 32 * k, EMB, 16 * (mi & 1), HID, 8 * nt, ncol, d, GINV, g, t);
                else flush_block(Q.dflat + f.W[3] + 32 * k, EMB + HID, 16 * (mi & 1), HID, 8 * nt, ncol, d, GINV, g, t);
            }
            bar_sync(bar2);
        }
        // ==================================================================== layers 0..4: h_{i+1}, then dW_{i+1} (or dWo for h_5)
#pragma unroll 1
        for (int i = 0; i < 5; ++i) {
            if (live) {
                if (i > 0) {
                    if (i == 3) {
#pragma unroll
                        for (int j = 0; j < 4; ++j)
#pragma unroll
                            for (int q = 0; q < 4; ++q) acc[j][q] = accS[j][q];
                    } else init_bias(acc, fs + F_b + i * HID, t);
#pragma unroll
                    for (int kk = 0; kk < 2; ++kk) {
                        AFrag<true> a;
                        afrag_from_c<true>(a, h[2 * kk], h[2 * kk + 1]);
                        kstep_w(acc, a, lo.f32 + I_WH + (i - 1) * 2 * PL32, RS32, kk);
                    }
                }
                relu_mask(h, acc);
                add_bias(h, fs + F_bc + i * HID, t);
#pragma unroll
                for (int kk = 0; kk < 2; ++kk) kstep_w(h, ca[kk], lo.f32 + I_FC + i * 2 * PL32, RS32, kk);     // + Fc_i c
                stage_x(xt, h, g, t);
            }
            bar_sync(bar1);
            if (i < 4) {   // gu_{i+1} (2 m-tiles) x h_{i+1} (4 n-tiles): one block per warp of the group
                const int mi = wl >> 2, nt = wl & 3;
                float d[4];
                consume<GW>(ga, xb, act, 0, 32 * (i + 1) + 16 * mi, nt, d);
                const int woff = i == 0 ? f.W[1] : i == 1 ? f.W[2] : i == 2 ? f.W[3] + EMB : f.W[4];
                flush_block(Q.dflat + woff, i == 2 ? EMB + HID : HID, 16 * mi, HID, 8 * nt, HID, d, GINV, g, t);
            } else {       // g_out x h_5 -> dWo: 4 n-tiles, each split over two warps (tiles 0-3 / 4-7)
                const int half = wl >> 2, nt = wl & 3;
                float d[4];
                consume<GW / 2>(ga, xb, act, 4 * half, C_GOUT, nt, d);
                flush_block(Q.dflat + f.Wo, HID, 0, 3, 8 * nt, HID, d, GINV, g, t);
            }
            bar_sync(bar2);
        }
        // ==================================================================== C phase: M_i = sum gu_{i+1}^T c, Mo = sum g_out^T c, bias sums
        if (live) {   // the split grid feature is still in the A fragments: (hi[0], hi[2]) of k-step kk = channels 8t + 4kk .. + 3 of row g
            *reinterpret_cast<uint4*>(xt + g * RS32 + 16 * t) = make_uint4(ca[0].hi[0], ca[0].hi[2], ca[1].hi[0], ca[1].hi[2]);
            *reinterpret_cast<uint4*>(xt + (g + 8) * RS32 + 16 * t) = make_uint4(ca[0].hi[1], ca[0].hi[3], ca[1].hi[1], ca[1].hi[3]);
            *reinterpret_cast<uint4*>(xt + X_PLANE + g * RS32 + 16 * t) = make_uint4(ca[0].lo[0], ca[0].lo[2], ca[1].lo[0], ca[1].lo[2]);
            *reinterpret_cast<uint4*>(xt + X_PLANE + (g + 8) * RS32 + 16 * t) = make_uint4(ca[0].lo[1], ca[0].lo[3], ca[1].lo[1], ca[1].lo[3]);
        }
        bar_sync(bar1);
#pragma unroll 1
        for (int blk = wl; blk < 47; blk += GW) {   // 8 m-tiles of gu_1..gu_4 + the g_out m-tile, x 4 n-tiles; then the 11 bias column sums
            float d[4];
            if (blk < 36) {
                const int mt = blk >> 2, nt = blk & 3;
                if (mt < 8) {
                    consume<GW>(ga, xb, act, 0, 32 + 16 * mt, nt, d);
                    flush_block(Q.scratch + SCR_M + (mt >> 1) * HID * HID, HID, 16 * (mt & 1), HID, 8 * nt, HID, d, GINV, g, t);
                } else {
                    consume<GW>(ga, xb, act, 0, C_GOUT, nt, d);
                    flush_block(Q.scratch + SCR_MO, HID, 0, 3, 8 * nt, HID, d, GINV, g, t);
                }
            } else {   // db_i = sum gu_i, dbo = sum g_out
                const int mt = blk - 36;
                consume_ones(ga, act, 16 * mt, d);
                if (t == 0) {
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {
                        const int col = 16 * mt + g + 8 * hh;
                        const int li = col >> 5;
                        const int boff = li == 0 ? f.b[0] : li == 1 ? f.b[1] : li == 2 ? f.b[2] : li == 3 ? f.b[3] : f.b[4];
                        if (col < 160) atomicAdd(Q.dflat + boff + (col & 31), d[2 * hh] * GINV);
                        else if (col < C_GOUT + 3) atomicAdd(Q.dflat + f.bo + (col - C_GOUT), d[2 * hh] * GINV);
                    }
                }
            }
        }
        bar_sync(bar2);
        // ==================================================================== Q phases: q = g_e cos(pB) per 32-feature chunk, dB = p^T q
#pragma unroll 1
        for (int k = 0; k < 3; ++k) {
            if (live) {
                float ge[4][4];
#pragma unroll
                for (int j = 0; j < 4; ++j) ge[j][0] = ge[j][1] = ge[j][2] = ge[j][3] = 0.0f;
#pragma unroll
                for (int src = 0; src < 2; ++src) {       // g_e = gu_0 W0 + gu_3 W3e; the A fragments come back from the G tile (single fp16)
#pragma unroll
                    for (int kk = 0; kk < 2; ++kk) {
                        AFrag<false> a;
                        ldsm4(gown + 2 * ((src ? 96 : 0) + 16 * kk), a.hi);
                        kstep_wt<false>(ge, a, lo.t96 + (src ? I_W3E : I_W0), RS96, kk, 4 * k);
                    }
                }
                // image column 32k + 8j + 2t + b holds feature 32k + 16 (j >> 1) + 4t + 2 (j & 1) + b; staged at natural feature columns
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int col = 16 * (j >> 1) + 4 * t + 2 * (j & 1), f0 = 32 * k + col;
                    const float2 B0 = *reinterpret_cast<const float2*>(fs + F_B + f0);
                    const float2 B1 = *reinterpret_cast<const float2*>(fs + F_B + EMBP + f0);
                    const float2 B2 = *reinterpret_cast<const float2*>(fs + F_B + 2 * EMBP + f0);
                    float sn, c00, c01, c10, c11;
                    ff_sincos(fmaf(p[0][2], B2.x, fmaf(p[0][1], B1.x, p[0][0] * B0.x)), sn, c00);
                    ff_sincos(fmaf(p[0][2], B2.y, fmaf(p[0][1], B1.y, p[0][0] * B0.y)), sn, c01);
                    ff_sincos(fmaf(p[1][2], B2.x, fmaf(p[1][1], B1.x, p[1][0] * B0.x)), sn, c10);
                    ff_sincos(fmaf(p[1][2], B2.y, fmaf(p[1][1], B1.y, p[1][0] * B0.y)), sn, c11);
                    uint32_t hi, lw;
                    split_f16(ge[j][0] * c00 * GINV, ge[j][1] * c01 * GINV, hi, lw);
                    *reinterpret_cast<uint32_t*>(xt + g * RS32 + 2 * col) = hi; *reinterpret_cast<uint32_t*>(xt + X_PLANE + g * RS32 + 2 * col) = lw;
                    split_f16(ge[j][2] * c10 * GINV, ge[j][3] * c11 * GINV, hi, lw);
                    *reinterpret_cast<uint32_t*>(xt + (g + 8) * RS32 + 2 * col) = hi; *reinterpret_cast<uint32_t*>(xt + X_PLANE + (g + 8) * RS32 + 2 * col) = lw;
                }
            }
            bar_sync(bar1);
            {   // rows 0..2 = p_hi^T q, rows 3..5 = p_lo^T q of the position m-tile: 4 n-tiles, each split over two warps
                const int half = wl >> 2, nt = wl & 3;
                float d[4];
                consume<GW / 2>(ga, xb, act, 4 * half, C_P, nt, d);
                const int c0 = 32 * k + 8 * nt + 2 * t;
                if (g < 6) {
                    float* dst = Q.dflat + f.B + (g % 3) * EMB + c0;
                    if (c0 < EMB) atomicAdd(dst, d[0]);
                    if (c0 + 1 < EMB) atomicAdd(dst + 1, d[1]);
                }
            }
            bar_sync(bar2);
        }
    }
}

// dFc_i = W_{i+1}^T M_i, dbc_i = W_{i+1}^T db_{i+1}  (i < 4; layer 3's hidden columns for i = 2),  dFc_4 = Wo^T Mo, dbc_4 = Wo^T dbo;
// clears the scratch for the next iteration.  One block.
__global__ void __launch_bounds__(1024) k_wgrad_finish(const float* __restrict__ flat, float* __restrict__ dflat, float* __restrict__ scratch) {
    const DecFlat f = DecFlat::make(32, 4);
    __shared__ float sM[SCR_N];
    for (int i = threadIdx.x; i < SCR_N; i += blockDim.x) { sM[i] = scratch[i]; scratch[i] = 0.0f; }
    __syncthreads();
    for (int idx = threadIdx.x; idx < 5 * HID * (HID + 1); idx += blockDim.x) {
        const int i = idx / (HID * (HID + 1)), rem = idx % (HID * (HID + 1)), m = rem / (HID + 1), ch = rem % (HID + 1);
        float acc = 0.0f;
        if (i < 4) {
            const float* W = flat + f.W[i + 1] + (i == 2 ? EMB : 0);
            const int ld = i == 2 ? EMB + HID : HID;
            for (int o = 0; o < HID; ++o) acc = fmaf(W[o * ld + m], ch < HID ? sM[SCR_M + (i * HID + o) * HID + ch] : dflat[f.b[i + 1] + o], acc);
        } else {
            for (int o = 0; o < 3; ++o) acc = fmaf(flat[f.Wo + o * HID + m], ch < HID ? sM[SCR_MO + o * HID + ch] : dflat[f.bo + o], acc);
        }
        if (ch < HID) atomicAdd(dflat + f.Fc[i] + m * 32 + ch, acc);
        else atomicAdd(dflat + f.bc[i] + m, acc);
    }
}

}  // namespace wgf

size_t wgrad_fused_img_bytes() { return wgf::IMG_BYTES; }
int wgrad_fused_scratch_floats() { return wgf::SCR_N; }

cudaError_t wgrad_fused_init() {
    static unsigned init = 0;
    int dev = 0; cudaGetDevice(&dev);
    if (!((init >> (dev & 31)) & 1u)) {
        const cudaError_t e = cudaFuncSetAttribute(wgf::k_wgrad_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, wgf::SMEM_BYTES);
        if (e != cudaSuccess) return e;
        init |= 1u << (dev & 31);
    }
    return cudaSuccess;
}
cudaError_t launch_build_wgimg(const float* flat, uint8_t* img, cudaStream_t st) {
    wgf::k_build_wgimg<<<8, 512, 0, st>>>(flat, img);
    return cudaGetLastError();
}
// P: the backward's decode parameters (rays, z, valid, g_raw, masks of the training forward; P.P samples, multiple of 16).
cudaError_t launch_wgrad_fused(const DecodeParams& P, const uint8_t* img, const float* flat, float* dflat, float* scratch, int n_sm, cudaStream_t st) {
    wgf::Params Q; Q.D = P; Q.img = img; Q.dflat = dflat; Q.scratch = scratch;
    const int nrounds = (P.P / TILE + wgf::GW - 1) / wgf::GW;
    const int grid = (nrounds + 1) / 2 < n_sm ? ((nrounds + 1) / 2 > 0 ? (nrounds + 1) / 2 : 1) : n_sm;
    wgf::k_wgrad_fused<<<grid, wgf::THREADS, wgf::SMEM_BYTES, st>>>(Q);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    wgf::k_wgrad_finish<<<1, 1024, 0, st>>>(flat, dflat, scratch);
    return cudaGetLastError();
}

}  // namespace nsb
