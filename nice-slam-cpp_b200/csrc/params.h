// Kernel parameter blocks shared between the host API (api.cu) and the kernels.
#pragma once
#include "common.cuh"

namespace nsb {

// Warps per CTA of the decoder kernels (one CTA per SM: a decoder's pre-split weights take 84-117 KB of shared memory).  The
// register file (64 K) divided by the thread count caps the registers per thread: 16 warps -> 128, 20 -> 96, 24 -> 80.
#ifndef NSB_FWD_WARPS
#define NSB_FWD_WARPS 16
#endif
#ifndef NSB_BWD_WARPS
#define NSB_BWD_WARPS 16
#endif
constexpr int FWD_WARPS = NSB_FWD_WARPS, FWD_THREADS = FWD_WARPS * 32;
constexpr int BWD_WARPS = NSB_BWD_WARPS, BWD_THREADS = BWD_WARPS * 32;

// Which decoders a launch evaluates and how the grid is split between them.
struct DecodeParams {
    const float* dec_flat[4];      // coarse, middle, fine, color flat parameter vectors (device)
    const float* wimg_fwd[4];      // pre-split shared-memory images of decoders 1..3 (k_build_wimg): forward orientation ...
    const float* wimg_bwd[4];      // ... and transposed for the backward kernel; a CTA copies its decoder's image with 16-byte loads
    const uint8_t* wimg_t5b[4];    // images of the tcgen05 backward (decode_bwd_t5.cu): transposed weight tiles, decoders 1 and 2
    const uint8_t* wimg_t5[4];     // images of the tcgen05 forward (decode_fwd_t5.cu): weight / bias tiles in the UMMA shared-memory layout
    const float* wimg_cmp[4];      // forward images with the grid-feature terms composed into the next layer (decoder_forward_composed)
    GridView grid[4];
    Bound bnd;
    // sample source: ray mode (rays + z) or points mode (pts != nullptr)
    const float* rays_o;           // [N][3]
    const float* rays_d;           // [N][3]
    const float* z;                // [N][S]
    const uint8_t* valid;          // [N] or nullptr: rays dropped by the inside filter are skipped
    const int* ray_list;           // tcgen05 forward: rays that pass the inside filter, compacted by k_zvals (nullptr: tiles walk all rays)
    int* ray_count;                // [0] its length, [1] CTAs finished (the last one clears both for the next launch)
    const float* pts;              // [P][3] or nullptr
    int S;                         // samples per ray (multiple of 16)
    int P;                         // total samples
    // outputs (forward): raw colour (P,4: r,g,b,-) and the three occupancy terms
    float* out_rgb;
    float* out_occ[3];             // coarse, middle, fine
    // grid partition: decoder d runs on CTAs [cta_begin[d], cta_begin[d+1])
    int cta_begin[5];
    // dynamic tile scheduling: the warps of decoder d draw tickets from tile_ctr[d] so that warps which meet many filtered rays,
    // or cheap tiles, simply take more tiles.  The counters are cleared by the kernel that precedes the decoder launch in the
    // stream (k_zvals / the composite kernels), which keeps the launch free of per-iteration host state (CUDA-graph replay).
    unsigned long long* tile_ctr;
    // ---- backward only ----
    const float* g_raw;            // [P][4] cotangent of raw (r,g,b,occ); occ already zero where out of bound
    int flags;                     // bit0 grid grads, bit1 colour-decoder weight grads (stash), bit2 ray grads
    float* d_rays;                 // [N][6]: d L / d rays_o, d L / d rays_d (atomically accumulated)
    float* stash;                  // [P][STASH_W] colour-decoder activations / gradients for the wgrad kernel
    uint32_t* masks;               // relu masks written by the training forward, read by the backward:
                                   //   mask_layout 0: [3][P/16][3][32] fragment-packed (mma.sync forward)
                                   //   mask_layout 1: [3][5][mask_stride] one 32-bit word per sample and layer (tcgen05 forward)
    int mask_layout, mask_stride;  // mask_layout: bit d set = decoder d uses layout 1 (0 = all fragment-packed, 0xE = all per-sample)
    const float* comp[4];          // composed weights of the tcgen05 forward (k_compose), per decoder
    unsigned long long* dbg;       // optional cycle counters of the tcgen05 forward (NSB_TC_TIMING builds), or nullptr
};

// Row layout of the colour-decoder weight-gradient stash (floats per sample).  The gradients at the block outputs (g_h) are NOT
// stashed: d Fc_i = W_{i+1}^T (sum_s g_u_{i+1} (x) c) and d bc_i = W_{i+1}^T d b_{i+1}, so k_wgrad contracts g_u with c instead and a tiny
// finishing kernel applies W^T (k_wgrad_finish) -- 552 instead of 712 floats per sample cross HBM three times.
namespace stash {
constexpr int E = 0;          // embedding e            96
constexpr int H = 96;         // h_1..h_5               5 x 32  (inputs of layers 1..4 and of the output layer)
constexpr int Cc = 256;       // grid feature c         32 (channel order)
constexpr int GU = 288;       // g_u_0..g_u_4           5 x 32  (gradient at the relu output, masked)
constexpr int GE = 448;       // g_e * cos(pB)          96
constexpr int GO = 544;       // g_out                  4
constexpr int Pp = 548;       // p                      4
constexpr int W = 552;        // = 8 (mod 32): conflict-free fragment loads from the ring in k_wgrad
}  // namespace stash

}  // namespace nsb
