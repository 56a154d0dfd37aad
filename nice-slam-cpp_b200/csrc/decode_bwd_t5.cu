// Backward decoder kernel on the 5th-generation tensor cores (tcgen05.mma kind::f16, operands in tensor memory): the data
// gradient of the occupancy decoders (middle, fine) for the mapping iteration -- d L / d grid features, scattered into the grid
// gradients -- i.e. the GRID-only instantiation of k_decode_bwd, which is a warp-MMA kernel pinned at T(HMMA) + T(ALU) (legacy HMMAs
// do not overlap with ALU work on sm_100a: decode_fwd_t5.cu, profiles/r2_microbench.json).  Same operand path as the forward:
//   warps 4g..4g+3 (tile group g)   one thread per sample = one TMEM lane.  Reads the relu masks the tcgen05 forward saved (one word per
//                                   sample and layer: exactly this thread's row), starts g_h5 = g_out Wo and g_c = g_out (Wo Fc_4) in
//                                   registers, then per layer: tcgen05.ld g_h -> mask -> fp16 split -> tcgen05.st -> arrive.
//   warps 16..19                    one elected lane per group issues  g_h = g_u W_i  (rewritten in place) and  g_c += g_u G_{i-1}
//                                   (G_l = W_{l+1} Fc_l: the composed grid-feature path of the forward) and commits to an mbarrier.
// Tensor-memory columns of a group:  g_h 32 | g_c 32 | x (hi 16 | lo 16): four groups of 128 columns.
// The tile's g_c then goes through the warp-aggregated scatter of decode_bwd.cuh (lane = channel, one whole-line reduction per
// distinct vertex), here over the 32 samples of a warp = two 16-sample ray segments back to back, with the trilinear setup done
// once per sample by its owner thread.
// This is the autograd backward of NICE::forward (NICE.cpp:43-50) for the stages whose loss has no colour term (Mapper.cpp:435-442).
#include "decode_bwd.cuh"
#include "params.h"
#include "t5_common.cuh"

namespace nsb {
namespace t5b {
using namespace t5;

constexpr int GH = 0, GC = 32, XC = 64, GCOLS = 128;   // tensor-memory columns inside a group
constexpr int WROW = 32;                                // samples a warp scatters in one walk
constexpr int SC_GC = 40;                               // row stride (floats) of the transposed gradient tile
constexpr int SC_W = WROW * SC_GC;                      // vertex weights by slot [32][8]
constexpr int SC_OFF = SC_W + WROW * 8;                 // vertex offsets by slot [32][8]
constexpr int SC_CELL = SC_OFF + WROW * 8;              // packed cell coordinates [32], then flush masks [32]
constexpr int SC_FLOATS = SC_CELL + 2 * WROW;           // per warp: 1856 floats

struct Smem {   // bytes; UMMA tiles on multiples of 1024 B, rows of 128 B = [hi 32 halves | lo 32 halves]
    static constexpr int WT = 0;                                   // 4 x [32 rows]: W_1^T, W_2^T, W_3^T (hidden columns), W_4^T  (row = input feature, k = output feature)
    static constexpr int GT = WT + 4 * 4096;                       // 4 x [32 rows]: G_0^T .. G_3^T restricted to the 32 channels that carry gradient (row = channel)
    static constexpr int WO = GT + 4 * 4096;                       // Wo[32] fp32 (occupancy output)
    static constexpr int WOC = WO + 32 * 4;                        // (Wo Fc_4)[32 channels] fp32
    static constexpr int IMG = (WOC + 32 * 4 + 1023) & ~1023;      // everything above: one decoder's backward image, prebuilt in global memory
    static constexpr int SCR = 2 * IMG;                            // both decoders' images are resident ([0] middle, [1] fine); then the scatter scratch, one block per compute warp
    static constexpr int BAR = SCR + (CTHREADS / 32) * SC_FLOATS * 4;   // mbarriers: full[NG], done[NG], image
    static constexpr int TMEMPTR = BAR + (2 * NG + 1) * 8;
    static constexpr int TICKET = TMEMPTR + 8;                     // [NG][4] ring of tiles
    static constexpr int TOTAL = TICKET + NG * 16;
};

// composed weights (global, per decoder), layout of k_compose: G[4][32][C] | bp[5][32] | woc[4][C] | boc[4]
template <int C>
__device__ void stage(uint8_t* sm, const float* __restrict__ flat, const float* __restrict__ comp, int tid, int nthr) {
    const DecFlat f = DecFlat::make(C, 1);
    for (int idx = tid; idx < 4 * 32 * 32; idx += nthr) {
        const int l = idx / 1024, m = (idx / 32) % 32, o = idx % 32;   // l = 0..3 <-> layers 1..4; element W_{l+1}[o][m]
        const float v = l == 2 ? flat[f.W[3] + o * (EMB + HID) + EMB + m] : flat[f.W[l + 1] + o * HID + m];
        put_w(sm + Smem::WT + l * 4096, m, o, v);
    }
    for (int idx = tid; idx < 4 * 32 * 32; idx += nthr) {
        const int l = idx / 1024, ch = (idx / 32) % 32, o = idx % 32;  // G_l[o][ch]
        put_w(sm + Smem::GT + l * 4096, ch, o, comp[(l * HID + o) * C + ch]);
    }
    float* wo = reinterpret_cast<float*>(sm + Smem::WO);
    float* woc = reinterpret_cast<float*>(sm + Smem::WOC);
    for (int i = tid; i < 32; i += nthr) { wo[i] = flat[f.Wo + i]; woc[i] = comp[4 * HID * C + 5 * HID + i]; }
}

struct ImgParams { const float* flat[4]; const float* comp[4]; uint8_t* img[4]; int mask; };
__global__ void __launch_bounds__(512) k_build_t5bimg(ImgParams W) {
    const int d = 1 + blockIdx.y;
    if (!((W.mask >> d) & 1)) return;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nthr = gridDim.x * blockDim.x;
    if (d == 1) stage<32>(W.img[1], W.flat[1], W.comp[1], tid, nthr);
    else stage<64>(W.img[2], W.flat[2], W.comp[2], tid, nthr);
}

// Scatter of the 32 samples of this warp (lane = sample for the setup, lane = channel for the walk): see scatter_tile in decode_bwd.cuh.
// scatter_stage leaves the warp's 32 gradient rows, vertex weights / offsets (by parity slot) and flush masks in the scratch;
// scatter_walk is the serial lane = channel pass over them (reads the scratch only, so the registers are free in between).
__device__ __forceinline__ void scatter_stage(const GridView& G, const Bound& bnd, const float (&p)[3], bool live, const float (&gc)[32],
                                              float* __restrict__ scr, int lane) {
    int* const scri = reinterpret_cast<int*>(scr);
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 8; ++i) *reinterpret_cast<float4*>(scr + lane * SC_GC + 4 * i) = make_float4(gc[4 * i], gc[4 * i + 1], gc[4 * i + 2], gc[4 * i + 3]);
    {
        Tri s;
        tri_setup(G, bnd, p, s);
        const int pc = (s.i0[0] & 1) | ((s.i0[1] & 1) << 1) | ((s.i0[2] & 1) << 2);
#pragma unroll
        for (int k = 0; k < 8; ++k) {                       // corner k of the cell -> slot k ^ pc (parity of the vertex coordinates)
            int off;
            const float w = tri_corner(G, s, k, off);
            scr[SC_W + 8 * lane + (k ^ pc)] = live ? w : 0.0f;
            scri[SC_OFF + 8 * lane + (k ^ pc)] = off * 4;       // byte offset (unsigned 32-bit: a grid is far below 4 GB)
        }
        scri[SC_CELL + lane] = s.i0[0] | (s.i0[1] << 10) | (s.i0[2] << 20);
    }
    __syncwarp();
    {
        uint32_t fm = 0;
        if (lane > 0) {
            const int a = scri[SC_CELL + lane - 1], b = scri[SC_CELL + lane];
            if (a != b) {
                const uint32_t keep = axis_keep(a & 1023, b & 1023, 0x55u, 0xaau) & axis_keep((a >> 10) & 1023, (b >> 10) & 1023, 0x33u, 0xccu) &
                                      axis_keep(a >> 20, b >> 20, 0x0fu, 0xf0u);
                fm = ~keep & 0xffu;
            }
        }
        scri[SC_CELL + WROW + lane] = (int)fm;
    }
    __syncwarp();
}
__device__ __forceinline__ void scatter_walk(const GridView& G, const float* __restrict__ scr, int lane) {
    const int* const scri = reinterpret_cast<const int*>(scr);
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.0f;
    const char* const base = reinterpret_cast<const char*>(G.grad + lane);   // + unsigned byte offset: two integer adds per address
#pragma unroll 1
    for (int s = 0; s <= WROW; ++s) {
        const uint32_t fm = s == WROW ? 0xffu : (uint32_t)scri[SC_CELL + WROW + s];
        if (fm) {                                            // warp-uniform: flush the vertices that leave with the previous sample's cell
            const int4 oa = *reinterpret_cast<const int4*>(scri + SC_OFF + 8 * (s - 1)), ob = *reinterpret_cast<const int4*>(scri + SC_OFF + 8 * (s - 1) + 4);
            const int off[8] = {oa.x, oa.y, oa.z, oa.w, ob.x, ob.y, ob.z, ob.w};
#define NSB_ADDR(j) reinterpret_cast<float*>(const_cast<char*>(base + (size_t)(uint32_t)off[j]))
#define NSB_FLUSH(j) { red_add_f32(NSB_ADDR(j), acc[j]); acc[j] = 0.0f; }
            // a move to a face neighbour (the common case) retires the four vertices of one face: six fixed patterns without per-slot tests
            switch (fm) {
                case 0x55u: NSB_FLUSH(0) NSB_FLUSH(2) NSB_FLUSH(4) NSB_FLUSH(6) break;
                case 0xaau: NSB_FLUSH(1) NSB_FLUSH(3) NSB_FLUSH(5) NSB_FLUSH(7) break;
                case 0x33u: NSB_FLUSH(0) NSB_FLUSH(1) NSB_FLUSH(4) NSB_FLUSH(5) break;
                case 0xccu: NSB_FLUSH(2) NSB_FLUSH(3) NSB_FLUSH(6) NSB_FLUSH(7) break;
                case 0x0fu: NSB_FLUSH(0) NSB_FLUSH(1) NSB_FLUSH(2) NSB_FLUSH(3) break;
                case 0xf0u: NSB_FLUSH(4) NSB_FLUSH(5) NSB_FLUSH(6) NSB_FLUSH(7) break;
                default:
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        if ((fm >> j) & 1u) { if (acc[j] != 0.0f) red_add_f32(NSB_ADDR(j), acc[j]); acc[j] = 0.0f; }
            }
#undef NSB_FLUSH
#undef NSB_ADDR
        }
        if (s == WROW) break;
        const float4 wa = *reinterpret_cast<const float4*>(scr + SC_W + 8 * s), wb = *reinterpret_cast<const float4*>(scr + SC_W + 8 * s + 4);
        const float v = scr[s * SC_GC + lane];
        acc[0] = fmaf(wa.x, v, acc[0]); acc[1] = fmaf(wa.y, v, acc[1]); acc[2] = fmaf(wa.z, v, acc[2]); acc[3] = fmaf(wa.w, v, acc[3]);
        acc[4] = fmaf(wb.x, v, acc[4]); acc[5] = fmaf(wb.y, v, acc[5]); acc[6] = fmaf(wb.z, v, acc[6]); acc[7] = fmaf(wb.w, v, acc[7]);
    }
}

__global__ void __launch_bounds__(THREADS, 1) k_decode_bwd_t5(const DecodeParams P) {
    extern __shared__ __align__(1024) uint8_t sm[];
    if (smem_u32(sm) & 1023u) __trap();
    using L = Smem;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ntiles1 = (P.P + TM - 1) / TM, ntiles = 2 * ntiles1;   // tickets
    const uint32_t bar0 = smem_u32(sm + L::BAR), bar_img = bar0 + 16 * NG;
    volatile uint32_t* tmem_ptr = reinterpret_cast<volatile uint32_t*>(sm + L::TMEMPTR);
    volatile int* ticket = reinterpret_cast<volatile int*>(sm + L::TICKET);
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(sm + L::TMEMPTR)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        for (int s = 0; s < NG; ++s) { mbar_init(bar0 + 8 * s, GTHREADS); mbar_init(bar0 + 8 * NG + 8 * s, 1); }
        mbar_init(bar_img, 1);
        asm volatile("fence.mbarrier_init.release.cluster;");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_img), "r"((uint32_t)(2 * L::IMG)) : "memory");
        for (int d = 0; d < 2; ++d) {
            const uint8_t* src = P.wimg_t5b[1 + d];
            for (int off = 0; off < L::IMG; off += 16384) {
                const uint32_t n = (uint32_t)(L::IMG - off < 16384 ? L::IMG - off : 16384);
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(smem_u32(sm + d * L::IMG + off)), "l"(src + off), "r"(n), "r"(bar_img) : "memory");
            }
        }
    }
    // one ticket counter for both decoders: tickets [0, ntiles) are the fine decoder's tiles (the more expensive ones first), [ntiles, 2 ntiles) the middle decoder's
    if ((tid & 127) == 0 && warp < NG * 4) {
        ticket[4 * (warp >> 2)] = (int)atomicAdd(P.tile_ctr + 1, 1ull);
        ticket[4 * (warp >> 2) + 1] = (int)atomicAdd(P.tile_ctr + 1, 1ull);
    }
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = *tmem_ptr;

    if (warp < NG * 4) {
        // ------------------------------------------------------------------ compute threads: one per sample (= TMEM lane)
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS_COMPUTE));
        const int grp = warp >> 2, wq = warp & 3, row = (wq << 5) | lane;
        const uint32_t tm = tmem + grp * GCOLS + ((uint32_t)(wq * 32) << 16);
        const uint32_t full = bar0 + 8 * grp, done = bar0 + 8 * NG + 8 * grp;
        float* scr = reinterpret_cast<float*>(sm + L::SCR) + warp * SC_FLOATS;
        // this thread's sample of a tile: cotangent of its occupancy output, relu masks of the training forward, position
        struct In { float gout; uint32_t m[5]; float p[3]; bool live; int dec; };
        auto load_in = [&](int tile, In& in) {
            const int dec = tile < ntiles1 ? 2 : 1;
            const int s = (tile < ntiles1 ? tile : tile - ntiles1) * TM + row;
            in.dec = dec;
            in.gout = 0.0f; in.live = false; in.p[0] = in.p[1] = in.p[2] = 0.0f;
#pragma unroll
            for (int i = 0; i < 5; ++i) in.m[i] = 0u;
            if (tile >= ntiles || s >= P.P) return;
            const int ray = s / P.S;
            const uint8_t ok = P.valid ? P.valid[ray] : (uint8_t)1;
            const float z = P.z[s];
            const float go = P.g_raw[4 * (size_t)s + 3];
            uint32_t mm[5];
#pragma unroll
            for (int i = 0; i < 5; ++i) mm[i] = P.masks[((size_t)(dec - 1) * 5 + i) * P.mask_stride + s];
            if (!ok) return;
            in.live = true; in.gout = go;
#pragma unroll
            for (int i = 0; i < 5; ++i) in.m[i] = mm[i];
#pragma unroll
            for (int a = 0; a < 3; ++a) in.p[a] = __fadd_rn(P.rays_o[3 * ray + a], __fmul_rn(P.rays_d[3 * ray + a], z));
        };
        uint32_t step = 0;
        // first step of a tile: the output layer in registers -- g_c = g_out (Wo Fc_4) starts the g_c accumulator, g_h5 = g_out Wo is
        // masked by layer 4's relu right away and goes out as the first operand
        auto first_step = [&](const In& in) {
            const float* wo = reinterpret_cast<const float*>(sm + (in.dec - 1) * L::IMG + L::WO);
            const float* woc = reinterpret_cast<const float*>(sm + (in.dec - 1) * L::IMG + L::WOC);
            float v[32];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float4 w4 = *reinterpret_cast<const float4*>(woc + 4 * i);
                v[4 * i] = in.gout * w4.x; v[4 * i + 1] = in.gout * w4.y; v[4 * i + 2] = in.gout * w4.z; v[4 * i + 3] = in.gout * w4.w;
            }
            tmem_st32(tm + GC, reinterpret_cast<const uint32_t*>(v));
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float4 w4 = *reinterpret_cast<const float4*>(wo + 4 * i);
                const float g4[4] = {in.gout * w4.x, in.gout * w4.y, in.gout * w4.z, in.gout * w4.w};
#pragma unroll
                for (int r = 0; r < 4; ++r) v[4 * i + r] = ((in.m[4] >> (4 * i + r)) & 1u) ? g4[r] : 0.0f;
            }
            store_operand(tm + XC, v);
            tmem_st_wait();
            fence_before();
            mbar_arrive(full);
            ++step;
        };
        int tile = ticket[4 * grp];
        In cur; load_in(tile, cur);
        mbar_wait(bar_img, 0);
        if (tile < ntiles) first_step(cur);
        for (int round = 0; tile < ntiles; ++round) {
            if (row == 0) ticket[4 * grp + (round + 2) % 3] = (int)atomicAdd(P.tile_ctr + 1, 1ull);   // visible after this round's group barrier
            const int tile_next = ticket[4 * grp + (round + 1) % 3];
            In nxt; load_in(tile_next, nxt);                 // the next tile's inputs travel while this tile's products run
            // ---- layers 3, 2, 1: read g_h, apply the layer's relu mask, hand g_u back
#pragma unroll 1
            for (int i = 3; i >= 1; --i) {
                mbar_wait(done, (step - 1) & 1);
                if (i == 3) group_sync(grp);                 // all 128 threads have just arrived for this product: the barrier costs nothing here
                fence_after();
                float v[32];
                tmem_ld32(tm + GH, v);
                tmem_ld_wait();
                const uint32_t m = cur.m[i];
#pragma unroll
                for (int k = 0; k < 32; ++k) v[k] = ((m >> k) & 1u) ? v[k] : 0.0f;
                store_operand(tm + XC, v);
                tmem_st_wait();
                fence_before();
                mbar_arrive(full);
                ++step;
            }
            // ---- g_c is complete after the layer-1 product: stage its scatter, start the next tile, then walk the scatter
            mbar_wait(done, (step - 1) & 1);
            fence_after();
            {
                float gc[32];
                tmem_ld32(tm + GC, gc);
                tmem_ld_wait();
                scatter_stage(P.grid[cur.dec], P.bnd, cur.p, cur.live, gc, scr, lane);
            }
            fence_before();
            if (tile_next < ntiles) first_step(nxt);         // its products run under the walk below
            scatter_walk(P.grid[cur.dec], scr, lane);
            tile = tile_next; cur = nxt;
        }
    } else {
        // ------------------------------------------------------------------ MMA issuer of one tile group (lane 0 issues)
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_ISSUE));
        const int grp = warp - NG * 4;
        const uint32_t tm = tmem + grp * GCOLS;
        const uint32_t full = bar0 + 8 * grp, done = bar0 + 8 * NG + 8 * grp;
        constexpr uint32_t I32 = make_idesc(128, 32);
        const uint64_t dbase = make_desc(smem_u32(sm));
        uint32_t step = 0;
        mbar_wait(bar_img, 0);
        for (int round = 0, tile = ticket[4 * grp]; tile < ntiles; ++round) {
            // g_u_{l+1} is in x:  g_h = g_u W_{l+1} (not needed below layer 1),  g_c += g_u G_l.  The group barrier sits right behind the
            // first product, where the compute threads take it too (they have all just arrived for that product).
#pragma unroll 1
            for (int l = 3; l >= 0; --l) {
                if (lane == 0) {
                    const int ib = (tile < ntiles1 ? 1 : 0) * L::IMG;       // fine decoder's image second
                    mbar_wait(full, step & 1); fence_after();
                    if (l > 0) issue3(tm + GH, tm + XC, dbase + (uint64_t)((ib + L::WT + l * 4096) >> 4), I32, true);
                    issue3(tm + GC, tm + XC, dbase + (uint64_t)((ib + L::GT + l * 4096) >> 4), I32, false);
                    mma_commit(done);
                }
                ++step;
                if (l == 3) { __syncwarp(); group_sync(grp); }
            }
            __syncwarp();
            tile = ticket[4 * grp + (round + 1) % 3];
        }
    }
    fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u));
}

}  // namespace t5b

size_t t5b_img_bytes() { return t5b::Smem::IMG; }
cudaError_t launch_build_t5bimg(const float* const flat[4], const float* const comp[4], uint8_t* const img[4], int mask, cudaStream_t st) {
    t5b::ImgParams W;
    for (int d = 0; d < 4; ++d) { W.flat[d] = flat[d]; W.comp[d] = comp[d]; W.img[d] = img[d]; }
    W.mask = mask & 0x6;
    if (!W.mask) return cudaSuccess;
    t5b::k_build_t5bimg<<<dim3(4, 2), 512, 0, st>>>(W);
    return cudaGetLastError();
}

// Decoders 1 and 2 (cta_begin[1] = 0 .. cta_begin[3] = grid), GRID-only; needs the per-sample relu masks of the tcgen05 forward.
cudaError_t launch_decode_bwd_t5(const DecodeParams& P, int grid, cudaStream_t st) {
    const size_t smem = (size_t)t5b::Smem::TOTAL;
    static unsigned attr_done = 0;
    int dev = 0; cudaGetDevice(&dev);
    if (!((attr_done >> (dev & 31)) & 1u)) {
        cudaError_t e = cudaFuncSetAttribute(t5b::k_decode_bwd_t5, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attr_done |= 1u << (dev & 31);
    }
    t5b::k_decode_bwd_t5<<<grid, t5::THREADS, smem, st>>>(P);
    return cudaGetLastError();
}

}  // namespace nsb
