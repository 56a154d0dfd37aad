// One (GRID, RAY, WG) instantiation of the backward decoder kernel per translation unit, so that the build can
// compile them in parallel (each takes ~1 min of ptxas).  NSB_BWD_COMBO selects the combination:
//   0: GRID          mapping, geometry stages        1: GRID|WG       mapping, colour stage
//   2: RAY           tracking                        3: GRID|RAY      bundle adjustment, fixed decoders
//   4: GRID|RAY|WG   bundle adjustment / full vjp    5: WG            decoder-only
#include "decode_bwd.cuh"

#ifndef NSB_BWD_COMBO
#error "define NSB_BWD_COMBO"
#endif

namespace nsb {
size_t decode_fwd_smem();

#if NSB_BWD_COMBO == 0
#define NSB_G true
#define NSB_R false
#define NSB_W false
#define NSB_NAME launch_decode_bwd_0
#elif NSB_BWD_COMBO == 1
#define NSB_G true
#define NSB_R false
#define NSB_W true
#define NSB_NAME launch_decode_bwd_1
#elif NSB_BWD_COMBO == 2
#define NSB_G false
#define NSB_R true
#define NSB_W false
#define NSB_NAME launch_decode_bwd_2
#elif NSB_BWD_COMBO == 3
#define NSB_G true
#define NSB_R true
#define NSB_W false
#define NSB_NAME launch_decode_bwd_3
#elif NSB_BWD_COMBO == 4
#define NSB_G true
#define NSB_R true
#define NSB_W true
#define NSB_NAME launch_decode_bwd_4
#else
#define NSB_G false
#define NSB_R false
#define NSB_W true
#define NSB_NAME launch_decode_bwd_5
#endif

template <bool P3>
static cudaError_t launch_one(const DecodeParams& P, int grid, cudaStream_t st) {
    const size_t smem = decode_fwd_smem() + sizeof(float) * BWD_WARPS * SCAT_FLOATS;
    cudaError_t e = cudaFuncSetAttribute(k_decode_bwd<P3, NSB_G, NSB_R, NSB_W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    k_decode_bwd<P3, NSB_G, NSB_R, NSB_W><<<grid, BWD_THREADS, smem, st>>>(P);
    return cudaGetLastError();
}

cudaError_t NSB_NAME(const DecodeParams& P, int precision, int grid, cudaStream_t st) {
    return precision == 0 ? launch_one<true>(P, grid, st) : launch_one<false>(P, grid, st);
}

}  // namespace nsb
