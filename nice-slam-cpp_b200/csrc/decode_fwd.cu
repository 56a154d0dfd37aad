// Forward decoder kernel: one launch evaluates every decoder the stage needs; the grid is partitioned between
// the decoders (each CTA keeps ONE decoder's weights resident in shared memory and walks that decoder's tiles).
#include "decode.cuh"
#include "decode_bwd.cuh"
#include "params.h"

namespace nsb {

#ifndef NSB_FWD_MIN_CTAS
#define NSB_FWD_MIN_CTAS 1   // 512 threads x 128 registers = the whole register file: one CTA of 16 warps per SM
#endif

// Positions of the thread's two samples (rows g and g+8 of the tile <-> samples base+2g, base+2g+1).
// Returns false when the whole tile is to be skipped (ray dropped by the inside filter).
__device__ __forceinline__ bool load_points(const DecodeParams& P, int base, int g, float (&p)[2][3], int (&sidx)[2]) {
#pragma unroll
    for (int r = 0; r < 2; ++r) sidx[r] = base + 2 * g + r;
    if (P.pts) {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int s = min(sidx[r], P.P - 1);
            p[r][0] = P.pts[3 * (size_t)s]; p[r][1] = P.pts[3 * (size_t)s + 1]; p[r][2] = P.pts[3 * (size_t)s + 2];
        }
        return true;
    }
    const int ray = base / P.S;
    // all loads are requested before the first use: one exposed latency instead of valid -> rays -> z
    const uint8_t ok = P.valid ? P.valid[ray] : (uint8_t)1;
    const float o[3] = {P.rays_o[3 * ray], P.rays_o[3 * ray + 1], P.rays_o[3 * ray + 2]};
    const float d[3] = {P.rays_d[3 * ray], P.rays_d[3 * ray + 1], P.rays_d[3 * ray + 2]};
    const float z[2] = {P.z[sidx[0]], P.z[sidx[1]]};
    if (!ok) return false;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
#pragma unroll
        for (int a = 0; a < 3; ++a) p[r][a] = __fadd_rn(o[a], __fmul_rn(d[a], z[r]));   // Renderer.cpp:121
    }
    return true;
}

// Packed relu masks of one tile: 3 words per lane (layers 0|1, 2|3, 4), [decoder-1][tile][3][32 lanes].
__device__ __forceinline__ void save_masks(const DecodeParams& P, int dec, int tile, int ntiles, int lane, const uint32_t (&m)[5]) {
    uint32_t* mb = P.masks + ((size_t)(dec - 1) * ntiles + tile) * 96 + lane;
    mb[0] = m[0] | (m[1] << 16); mb[32] = m[2] | (m[3] << 16); mb[64] = m[4];
}

// TRAIN: the launch is the forward half of a training step -- relu masks are kept for the backward kernel (which then
// needs no forward recomputation) and, when P.stash is set, the colour decoder's activations go to the wgrad stash.
template <bool P3, bool TRAIN>
__global__ void __launch_bounds__(FWD_THREADS, NSB_FWD_MIN_CTAS) k_decode_fwd(const DecodeParams P) {
    extern __shared__ __align__(128) float sm[];
    int dec = 0;
#pragma unroll
    for (int d = 1; d < 4; ++d) if ((int)blockIdx.x >= P.cta_begin[d]) dec = d;
    const int cta = blockIdx.x - P.cta_begin[dec], ncta = P.cta_begin[dec + 1] - P.cta_begin[dec];
    if (dec == 0) stage_coarse(sm, P.dec_flat[0], threadIdx.x, blockDim.x);
    // decoders 1 / 2 and the colour decoder without a stash run on the composed image (decoder_forward_composed)
    else if (dec == 2) load_decoder_image<64>(sm, P.wimg_cmp[2], threadIdx.x, blockDim.x);
    else load_decoder_image<32>(sm, (dec == 3 && TRAIN && P.stash) ? P.wimg_fwd[3] : P.wimg_cmp[dec], threadIdx.x, blockDim.x);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int ntiles = (P.P + TILE - 1) / TILE;
    (void)cta; (void)ncta; (void)warp;
    TileQueue q; q.init(P.tile_ctr + dec, 0ull, ntiles, lane);
    for (int tile = q.next(lane); tile >= 0; tile = q.next(lane)) {
        float p[2][3]; int sidx[2];
        if (!load_points(P, tile * TILE, g, p, sidx)) continue;
        if (dec == 0) {
            float c[2][8], out[2]; uint32_t masks[5];
            gather8(P.grid[0], P.bnd, p[0], t, c[0]);
            gather8(P.grid[0], P.bnd, p[1], t, c[1]);
            coarse_forward<P3>(sm, c, g, t, out, masks);
            if (TRAIN) save_masks(P, 1, tile, ntiles, lane, masks);     // the coarse mapper runs no other decoder: slot of decoder 1
            if (t == 0) {
#pragma unroll
                for (int r = 0; r < 2; ++r) if (sidx[r] < P.P) P.out_occ[0][sidx[r]] = out[r];
            }
        } else if (dec == 1 || dec == 3) {
            float c[2][8], out[2][4], h[4][4]; uint32_t masks[5];
            gather8(P.grid[dec], P.bnd, p[0], t, c[0]);
            gather8(P.grid[dec], P.bnd, p[1], t, c[1]);
            if (dec == 1) {
                decoder_forward_composed<32, 1, P3>(sm, p, c, g, t, out, masks);
                if (TRAIN) save_masks(P, 1, tile, ntiles, lane, masks);
                if (t == 0) {
#pragma unroll
                    for (int r = 0; r < 2; ++r) if (sidx[r] < P.P) P.out_occ[1][sidx[r]] = out[r][0];
                }
            } else {
                if (TRAIN && P.stash) {
                    float* st0 = P.stash + (size_t)sidx[0] * stash::W; float* st1 = P.stash + (size_t)sidx[1] * stash::W;
                    *reinterpret_cast<float4*>(st0 + stash::Cc + 8 * t) = make_float4(c[0][0], c[0][1], c[0][2], c[0][3]);
                    *reinterpret_cast<float4*>(st0 + stash::Cc + 8 * t + 4) = make_float4(c[0][4], c[0][5], c[0][6], c[0][7]);
                    *reinterpret_cast<float4*>(st1 + stash::Cc + 8 * t) = make_float4(c[1][0], c[1][1], c[1][2], c[1][3]);
                    *reinterpret_cast<float4*>(st1 + stash::Cc + 8 * t + 4) = make_float4(c[1][4], c[1][5], c[1][6], c[1][7]);
                    decoder_forward<32, 4, P3, TRAIN>(sm, p, c, g, t, out, masks, h, st0, st1);
                } else {
                    decoder_forward_composed<32, 4, P3>(sm, p, c, g, t, out, masks);
                }
                if (TRAIN) save_masks(P, 3, tile, ntiles, lane, masks);
                if (t == 0) {
#pragma unroll
                    for (int r = 0; r < 2; ++r)
                        if (sidx[r] < P.P) *reinterpret_cast<float4*>(P.out_rgb + 4 * (size_t)sidx[r]) = make_float4(out[r][0], out[r][1], out[r][2], 0.0f);
                }
            }
        } else {
            float c[2][16], out[2][4], h[4][4]; uint32_t masks[5];
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                gather8(P.grid[2], P.bnd, p[r], t, c[r]);          // fine features ...
                gather8(P.grid[1], P.bnd, p[r], t, c[r] + 8);      // ... cat middle features (MLP.cpp:79-84)
            }
            decoder_forward_composed<64, 1, P3>(sm, p, c, g, t, out, masks);
            if (TRAIN) save_masks(P, 2, tile, ntiles, lane, masks);
            if (t == 0) {
#pragma unroll
                for (int r = 0; r < 2; ++r) if (sidx[r] < P.P) P.out_occ[2][sidx[r]] = out[r][0];
            }
        }
    }
}

// Grid sampling in isolation (the K6 row of SURVEY 2.1): the same quad-cooperative trilinear gather as the fused kernels
// for the middle, fine and colour grids of every sample, features reduced to one float per sample so that nothing but the
// gather is timed.  Used by nsb_bench_gather for the "grid sampling vs L2/HBM roofline" number.
__global__ void __launch_bounds__(256) k_gather_only(const DecodeParams P, float* __restrict__ out) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int nwarps = (gridDim.x * blockDim.x) >> 5, ntiles = P.P / TILE;
    for (int tile = warp; tile < ntiles; tile += nwarps) {
        float p[2][3]; int sidx[2];
        if (!load_points(P, tile * TILE, g, p, sidx)) continue;
        float acc[2] = {0.0f, 0.0f};
#pragma unroll
        for (int lv = 1; lv < 4; ++lv) {
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                float c[8];
                gather8(P.grid[lv], P.bnd, p[r], t, c);
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[r] += c[i];
            }
        }
        acc[0] = quad_sum(acc[0]); acc[1] = quad_sum(acc[1]);
        if (t == 0) { out[sidx[0]] = acc[0]; out[sidx[1]] = acc[1]; }
    }
}

// Backward of the coarse stage (the coarse mapper, Mapper.cpp:335-338,351-352: only grid_coarse is optimised): data gradient of
// MLP_no_xyz from the saved relu masks, scattered into the coarse grid's gradient with the same quad-wide vector reductions as the
// other levels.  Warp = 16-sample tile, static striding (the launch is small: 32 samples per ray, one decoder).
template <bool P3>
__global__ void __launch_bounds__(256) k_coarse_bwd(const DecodeParams P) {
    extern __shared__ __align__(128) float sm[];
    stage_coarse_bwd(sm, P.dec_flat[0], threadIdx.x, blockDim.x);
    __syncthreads();
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int nwarps = (gridDim.x * blockDim.x) >> 5, ntiles = P.P / TILE;
    for (int tile = warp; tile < ntiles; tile += nwarps) {
        float p[2][3]; int sidx[2];
        const uint32_t* mb = P.masks + (size_t)tile * 96 + lane;
        const uint32_t m0 = mb[0], m1 = mb[32], m2 = mb[64];
        if (!load_points(P, tile * TILE, g, p, sidx)) continue;
        const float gout[2] = {P.g_raw[4 * (size_t)sidx[0] + 3], P.g_raw[4 * (size_t)sidx[1] + 3]};
        if (!__any_sync(0xffffffffu, gout[0] != 0.0f || gout[1] != 0.0f)) continue;
        const uint32_t masks[5] = {m0 & 0xffffu, m0 >> 16, m1 & 0xffffu, m1 >> 16, m2};
        float gcf[2][8], gp[2][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};
        coarse_backward<P3>(sm, g, t, gout, masks, gcf);
        grid_backward<true, false>(P.grid[0], P.bnd, p, gcf, t, gp);
    }
}
cudaError_t launch_coarse_bwd(const DecodeParams& P, int precision, int grid, cudaStream_t st) {
    const size_t smem = sizeof(float) * CoarseSmem::TOTAL;
    cudaError_t e = precision == 0 ? cudaFuncSetAttribute(k_coarse_bwd<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                                   : cudaFuncSetAttribute(k_coarse_bwd<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    if (precision == 0) k_coarse_bwd<true><<<grid, 256, smem, st>>>(P);
    else k_coarse_bwd<false><<<grid, 256, smem, st>>>(P);
    return cudaGetLastError();
}

cudaError_t launch_gather_only(const DecodeParams& P, float* out, int grid, cudaStream_t st) {
    k_gather_only<<<grid, 256, 0, st>>>(P, out);
    return cudaGetLastError();
}

// Builds the pre-split shared-memory images of the decoders in `mask` (bit d): blockIdx.y = 3 (d - 1) + kind, kind 0 = forward
// (plain, only the colour decoder needs it: stash iterations), 1 = backward (transposed), 2 = forward composed (reads k_compose's output).
struct WimgParams { const float* flat[4]; const float* comp[4]; float* fwd[4]; float* bwd[4]; float* cmp[4]; int mask, cmp_mask; };
__global__ void __launch_bounds__(512) k_build_wimg(WimgParams W) {
    const int d = 1 + blockIdx.y / 3, kind = blockIdx.y % 3;
    if (!(((kind == 2 ? W.cmp_mask : W.mask) >> d) & 1)) return;
    if (kind == 0 && d != 3) return;
    if (kind == 1 && blockIdx.x < 4) {   // the composed G_l^T slots of the backward image: one block each
        if (d == 1) stage_backward_G<32, 1>(W.bwd[1], W.flat[1], blockIdx.x);
        else if (d == 2) stage_backward_G<64, 1>(W.bwd[2], W.flat[2], blockIdx.x);
        else stage_backward_G<32, 4>(W.bwd[3], W.flat[3], blockIdx.x);
    }
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nthr = gridDim.x * blockDim.x;
    if (d == 1) { if (kind == 1) stage_decoder<32, 1, true>(W.bwd[1], W.flat[1], tid, nthr); else stage_decoder_composed<32, 1>(W.cmp[1], W.flat[1], W.comp[1], tid, nthr); }
    else if (d == 2) { if (kind == 1) stage_decoder<64, 1, true>(W.bwd[2], W.flat[2], tid, nthr); else stage_decoder_composed<64, 1>(W.cmp[2], W.flat[2], W.comp[2], tid, nthr); }
    else {
        if (kind == 0) stage_decoder<32, 4, false>(W.fwd[3], W.flat[3], tid, nthr);
        else if (kind == 1) stage_decoder<32, 4, true>(W.bwd[3], W.flat[3], tid, nthr);
        else stage_decoder_composed<32, 4>(W.cmp[3], W.flat[3], W.comp[3], tid, nthr);
    }
}
cudaError_t launch_build_wimg(const float* const flat[4], const float* const comp[4], float* const img_fwd[4], float* const img_bwd[4], float* const img_cmp[4],
                              int mask, int cmp_mask, cudaStream_t st) {
    WimgParams W;
    for (int d = 0; d < 4; ++d) { W.flat[d] = flat[d]; W.comp[d] = comp[d]; W.fwd[d] = img_fwd[d]; W.bwd[d] = img_bwd[d]; W.cmp[d] = img_cmp[d]; }
    W.mask = mask; W.cmp_mask = cmp_mask;
    k_build_wimg<<<dim3(24, 9), 512, 0, st>>>(W);
    return cudaGetLastError();
}
size_t wimg_floats(int which) { return which == 2 ? DecSmem<64>::TOTAL : DecSmem<32>::TOTAL; }

size_t decode_fwd_smem() { return sizeof(float) * DecSmem<64>::TOTAL; }

template <bool P3, bool TRAIN>
static cudaError_t launch_fwd_one(const DecodeParams& P, int grid, cudaStream_t st) {
    const size_t smem = decode_fwd_smem();
    cudaError_t e = cudaFuncSetAttribute(k_decode_fwd<P3, TRAIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    k_decode_fwd<P3, TRAIN><<<grid, FWD_THREADS, smem, st>>>(P);
    return cudaGetLastError();
}

// P.masks != nullptr selects the training variant (relu masks saved, optional colour-decoder stash).
cudaError_t launch_decode_fwd(const DecodeParams& P, int precision, int grid, cudaStream_t st) {
    if (P.masks) return precision == 0 ? launch_fwd_one<true, true>(P, grid, st) : launch_fwd_one<false, true>(P, grid, st);
    return precision == 0 ? launch_fwd_one<true, false>(P, grid, st) : launch_fwd_one<false, false>(P, grid, st);
}

int decode_fwd_occupancy(int precision) {
    int nb = 0;
    const size_t smem = decode_fwd_smem();
    cudaFuncSetAttribute(k_decode_fwd<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k_decode_fwd<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (precision == 0) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_decode_fwd<true, true>, FWD_THREADS, smem);
    else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_decode_fwd<false, true>, FWD_THREADS, smem);
    return nb;
}

}  // namespace nsb
