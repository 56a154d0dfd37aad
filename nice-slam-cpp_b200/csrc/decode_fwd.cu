// Forward decoder kernel: one launch evaluates every decoder the stage needs; the grid is partitioned between
// the decoders (each CTA keeps ONE decoder's weights resident in shared memory and walks that decoder's tiles).
#include "decode.cuh"
#include "params.h"

namespace nsb {

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
    if (P.valid && !P.valid[ray]) return false;
    const float o[3] = {P.rays_o[3 * ray], P.rays_o[3 * ray + 1], P.rays_o[3 * ray + 2]};
    const float d[3] = {P.rays_d[3 * ray], P.rays_d[3 * ray + 1], P.rays_d[3 * ray + 2]};
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const float z = P.z[sidx[r]];
#pragma unroll
        for (int a = 0; a < 3; ++a) p[r][a] = __fadd_rn(o[a], __fmul_rn(d[a], z));   // Renderer.cpp:121
    }
    return true;
}

template <bool P3>
__global__ void __launch_bounds__(DECODE_THREADS) k_decode_fwd(const DecodeParams P) {
    extern __shared__ __align__(128) float sm[];
    int dec = 0;
#pragma unroll
    for (int d = 1; d < 4; ++d) if ((int)blockIdx.x >= P.cta_begin[d]) dec = d;
    const int cta = blockIdx.x - P.cta_begin[dec], ncta = P.cta_begin[dec + 1] - P.cta_begin[dec];
    if (dec == 0) stage_coarse(sm, P.dec_flat[0], threadIdx.x, blockDim.x);
    else if (dec == 1) stage_decoder<32, 1>(sm, P.dec_flat[1], threadIdx.x, blockDim.x);
    else if (dec == 2) stage_decoder<64, 1>(sm, P.dec_flat[2], threadIdx.x, blockDim.x);
    else stage_decoder<32, 4>(sm, P.dec_flat[3], threadIdx.x, blockDim.x);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int ntiles = (P.P + TILE - 1) / TILE;
    for (int tile = cta * DECODE_WARPS + warp; tile < ntiles; tile += ncta * DECODE_WARPS) {
        float p[2][3]; int sidx[2];
        if (!load_points(P, tile * TILE, g, p, sidx)) continue;
        if (dec == 0) {
            float c[2][8], out[2];
            gather8(P.grid[0], P.bnd, p[0], t, c[0]);
            gather8(P.grid[0], P.bnd, p[1], t, c[1]);
            coarse_forward<P3>(sm, c, g, t, out);
            if (t == 0) {
#pragma unroll
                for (int r = 0; r < 2; ++r) if (sidx[r] < P.P) P.out_occ[0][sidx[r]] = out[r];
            }
        } else if (dec == 1 || dec == 3) {
            float c[2][8], out[2][4], h[4][4]; uint32_t masks[5];
            gather8(P.grid[dec], P.bnd, p[0], t, c[0]);
            gather8(P.grid[dec], P.bnd, p[1], t, c[1]);
            if (dec == 1) {
                decoder_forward<32, 1, P3, false>(sm, p, c, g, t, out, masks, h, nullptr, nullptr);
                if (t == 0) {
#pragma unroll
                    for (int r = 0; r < 2; ++r) if (sidx[r] < P.P) P.out_occ[1][sidx[r]] = out[r][0];
                }
            } else {
                decoder_forward<32, 4, P3, false>(sm, p, c, g, t, out, masks, h, nullptr, nullptr);
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
            decoder_forward<64, 1, P3, false>(sm, p, c, g, t, out, masks, h, nullptr, nullptr);
            if (t == 0) {
#pragma unroll
                for (int r = 0; r < 2; ++r) if (sidx[r] < P.P) P.out_occ[2][sidx[r]] = out[r][0];
            }
        }
    }
}

size_t decode_fwd_smem() { return sizeof(float) * DecSmem<64>::TOTAL; }

cudaError_t launch_decode_fwd(const DecodeParams& P, int precision, int grid, cudaStream_t st) {
    const size_t smem = decode_fwd_smem();
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(k_decode_fwd<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(k_decode_fwd<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attr_done = true;
    }
    if (precision == 0) k_decode_fwd<true><<<grid, DECODE_THREADS, smem, st>>>(P);
    else k_decode_fwd<false><<<grid, DECODE_THREADS, smem, st>>>(P);
    return cudaGetLastError();
}

int decode_fwd_occupancy(int precision) {
    int nb = 0;
    const size_t smem = decode_fwd_smem();
    cudaFuncSetAttribute(k_decode_fwd<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k_decode_fwd<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (precision == 0) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_decode_fwd<true>, DECODE_THREADS, smem);
    else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_decode_fwd<false>, DECODE_THREADS, smem);
    return nb;
}

}  // namespace nsb
