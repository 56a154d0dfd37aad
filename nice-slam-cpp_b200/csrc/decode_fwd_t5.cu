// Forward decoder kernel on the 5th-generation tensor cores, third generation: tcgen05.mma kind::f16 with EVERY A operand in
// TENSOR MEMORY (grid features, Fourier features, hidden activations), weights (B operands) resident in shared memory, fp32
// accumulators in tensor memory.  Same fp32-grade product as the warp-MMA path,
//     D += a_lo.b_hi + a_hi.b_lo + a_hi.b_hi,   x = hi + lo,  hi = fp16(x), lo = fp16(x - hi).
//
// Why this exists (measured, tools/microbench, profiles/r2_microbench.json): on sm_100a a legacy warp-level HMMA does not overlap
// with FP32/ALU work of the same SM sub-partition -- 8 HMMA + 64 FFMA per warp take 411 clk where either alone takes 280 / 265, and
// HMMA-only warps next to FFMA-only warps take the SUM (555 clk).  k_decode_fwd is therefore pinned at T(HMMA) + T(ALU); only
// the asynchronous tcgen05 pipe runs the products underneath the per-value work (sines, fp16 splits, relu, gather).
//
// One CTA per SM keeps ONE decoder's weights resident and runs FOUR tiles of 128 samples concurrently (three for the fine decoder, whose
// 64 grid channels need 160 of the 512 tensor-memory columns per tile):
//   warps 4g..4g+3 (tile group g)   one thread per sample = one TMEM lane.  Quad-cooperative trilinear gather (a quad fetches a voxel
//                                   line in one wavefront) handed to the owner lane through a per-warp scratch; Fourier features;
//                                   per-layer epilogues (tcgen05.ld, bias, relu, mask, fp16 split).  Operands go back to tensor
//                                   memory with tcgen05.st as packed f16x2 words: no swizzled shared-memory stores, no bank
//                                   conflicts, no async-proxy fence.
//   warps 16..19                    one elected lane per group issues the tcgen05.mma stream and commits to an mbarrier
// The four issuer warps give registers back (setmaxnreg.dec 32) and the sixteen compute warps take them (setmaxnreg.inc 120).
// Tensor-memory columns of a group (64 + C + 32 of the 512):  acc0 | accS | c (hi 16 | lo 16 per 32 channels) | x (hi 16 | lo 16).
// Biases are products too: a constant shared-memory operand (1, 1, 0, ..) times a tile holding (hi(b'), lo(b'), 0, ..) per output
// adds b' to the accumulator inside the tensor pipe, so the epilogue is relu + sign-bit mask + split only.
// The decoder's image (weight / bias / ones tiles, small tables; built once per weight update) arrives by ONE-dimensional bulk TMA
// (cp.async.bulk -> mbarrier complete_tx) while the first tile's gather is already running; tiles are drawn from a ticket counter.
// Algebra (k_compose): a_{i+1} = W_{i+1} relu(a_i) + G_i c + b'_{i+1}, G_i = W_{i+1} Fc_i.  c stays in its columns for the whole
// tile and its term joins each layer's product, so a layer's accumulator is rewritten in place (acc0 serves layers 0, 1, 2, 4;
// accS collects the skip layer) -- 64 accumulator columns per tile instead of the 160 of decode_fwd_tc16.cu.
// Replaces NICE::forward / MLP::forward (NICE.cpp:16-51, MLP.cpp:76-102); selected with NSB_TCGEN05=3 when no wgrad stash is needed.
#include "decode.cuh"
#include "params.h"
#include "t5_common.cuh"

namespace nsb {
namespace t5 {

constexpr int ACC0 = 0, ACCS = 32, CCOL = 64;  // tensor-memory columns inside a group
__host__ __device__ constexpr int xcol(int C) { return CCOL + C; }
__host__ __device__ constexpr int gcols(int C) { return CCOL + C + 32; }
__host__ __device__ constexpr int ngroups(int C) { return 512 / gcols(C) < NG ? 512 / gcols(C) : NG; }
constexpr int SCR_ROW = 36;                    // floats per row of the gather scratch (conflict-free 16-byte stores and loads)

// composed weights (global, per decoder), layout of k_compose: G[4][32][C] | bp[5][32] | woc[4][C] | boc[4]
__host__ __device__ constexpr int comp_bp(int C) { return 4 * HID * C; }
__host__ __device__ constexpr int comp_woc(int C) { return comp_bp(C) + 5 * HID; }
__host__ __device__ constexpr int comp_boc(int C) { return comp_woc(C) + 4 * C; }

template <int C>
struct Smem {   // bytes; every UMMA tile starts on a multiple of 1024 B, rows are 128 B = [hi 32 halves | lo 32 halves]
    static constexpr int WE = 0;                                   // 3 chunks x [64 rows]: rows 0-31 W0, 32-63 W3 (embedding columns)
    static constexpr int WH = WE + 3 * 64 * 128;                   // 4 x [32 rows]: W1, W2, W3 (hidden columns), W4
    static constexpr int G = WH + 4 * 32 * 128;                    // 4 x (C/32) x [32 rows]: G_i restricted to channel chunk cc at (i * C/32 + cc)
    static constexpr int BT = G + 4 * (C / 32) * 32 * 128;         // bias tile [64 rows]: k16 step 0 = b'_0 (rows 0-31) | b'_3 (rows 32-63), steps 1, 2, 3 = b'_1, b'_2, b'_4
    static constexpr int ONE = BT + 64 * 128;                      // constant A operand of the bias products [128 rows]: halves 0, 1 = 1.0, rest 0
    static constexpr int BM = ONE + 128 * 128;                     // Fourier matrix [3][96] fp32
    static constexpr int BP = BM + 3 * EMBP * 4;                   // b'[5][32]
    static constexpr int WO = BP + 5 * HID * 4;                    // Wo[4][32]
    static constexpr int WOC = WO + 4 * HID * 4;                   // (Wo Fc_4)[4][C]
    static constexpr int BOC = WOC + 4 * C * 4;                    // const[4]
    static constexpr int IMG = BOC + 16;                           // everything above is the decoder's image, prebuilt in global memory (k_build_t5img)
    static constexpr int SCR = IMG;                                // gather scratch: C/32 blocks of [32][SCR_ROW] fp32 per active compute warp
    static constexpr int SCRW = (C / 32) * 32 * SCR_ROW * 4;       // bytes per warp
    static constexpr int STX = SCR + ngroups(C) * 4 * SCRW;        // stash transposition block [32][SCR_ROW] per compute warp (32-channel decoders only: the colour decoder)
    static constexpr int STXW = C == 32 ? 32 * SCR_ROW * 4 : 0;
    static constexpr int BAR = STX + ngroups(C) * 4 * STXW;        // mbarriers: full[NG], done[NG], image
    static constexpr int TMEMPTR = BAR + (3 * NG + 1) * 8;         // full[NG], done[NG], image, ready[NG]
    static constexpr int TICKET = TMEMPTR + 8;                     // [NG][3] ring of tiles drawn by a group's row 0 (current, next, the one after)
    static constexpr int TOTAL = TICKET + NG * 16;
};

template <int C, int O>
__device__ void stage(uint8_t* sm, const float* __restrict__ flat, const float* __restrict__ comp, int tid, int nthr) {
    using L = Smem<C>;
    const DecFlat f = DecFlat::make(C, O);
    for (int idx = tid; idx < 3 * 64 * 32; idx += nthr) {
        const int j = idx / 2048, r = (idx / 32) % 64, k = idx % 32, ft = 32 * j + k;
        float v = 0.0f;
        if (ft < EMB) v = r < 32 ? flat[f.W[0] + r * EMB + ft] : flat[f.W[3] + (r - 32) * (EMB + HID) + ft];
        put_w(sm + L::WE + j * 8192, r, k, v);
    }
    for (int idx = tid; idx < 4 * 32 * 32; idx += nthr) {
        const int l = idx / 1024, o = (idx / 32) % 32, k = idx % 32;   // l = 0..3 <-> layers 1..4
        const float v = l == 2 ? flat[f.W[3] + o * (EMB + HID) + EMB + k] : flat[f.W[l + 1] + o * HID + k];
        put_w(sm + L::WH + l * 4096, o, k, v);
    }
    for (int idx = tid; idx < 4 * (C / 32) * 32 * 32; idx += nthr) {
        const int tile = idx / 1024, i = tile / (C / 32), cc = tile % (C / 32), o = (idx / 32) % 32, k = idx % 32;
        put_w(sm + L::G + tile * 4096, o, k, comp[(i * HID + o) * C + 32 * cc + k]);
    }
    for (int idx = tid; idx < 64 * 64; idx += nthr) {               // bias tile: half 16 s of row r = hi(b), half 16 s + 1 = lo(b), rest 0
        const int r = idx / 64, hpos = idx % 64, st = hpos / 16, kk = hpos % 16;
        float b = 0.0f;
        if (st == 0) b = comp[comp_bp(C) + (r < 32 ? 0 : 3) * HID + (r & 31)];
        else if (r < 32) b = comp[comp_bp(C) + (st == 3 ? 4 : st) * HID + r];
        const __half h = __float2half_rn(b);
        const __half l = __float2half_rn(b - __half2float(h));
        *reinterpret_cast<__half*>(sm + L::BT + chunk_off(r, hpos >> 3) + (hpos & 7) * 2) = kk == 0 ? h : kk == 1 ? l : __float2half_rn(0.0f);
    }
    for (int idx = tid; idx < 128 * 32; idx += nthr) {              // ones operand: first word of every row = (1.0h, 1.0h)
        const int r = idx / 32, wd = idx % 32;
        *reinterpret_cast<uint32_t*>(sm + L::ONE + chunk_off(r, wd >> 2) + (wd & 3) * 4) = wd == 0 ? 0x3C003C00u : 0u;
    }
    float* bm = reinterpret_cast<float*>(sm + L::BM);
    for (int i = tid; i < 3 * EMBP; i += nthr) { const int d = i / EMBP, c = i % EMBP; bm[i] = c < EMB ? flat[f.B + d * EMB + c] : 0.0f; }
    float* bp = reinterpret_cast<float*>(sm + L::BP);
    for (int i = tid; i < 5 * HID; i += nthr) bp[i] = comp[comp_bp(C) + i];
    float* wo = reinterpret_cast<float*>(sm + L::WO);
    for (int i = tid; i < 4 * HID; i += nthr) wo[i] = (i / HID) < O ? flat[f.Wo + i] : 0.0f;
    float* woc = reinterpret_cast<float*>(sm + L::WOC);
    for (int i = tid; i < 4 * C; i += nthr) woc[i] = comp[comp_woc(C) + i];
    float* boc = reinterpret_cast<float*>(sm + L::BOC);
    if (tid < 4) boc[tid] = comp[comp_boc(C) + tid];
}

// 32 values per thread (its own sample's row) -> columns [col, col + 32) of the samples' stash rows, transposed through a per-warp block so
// that every store instruction writes whole 128-byte pieces of four rows (row-per-thread stores touch 32 lines with 16 bytes each and
// cost 0.1 ms per launch).  srow: the thread's sample index or -1.
__device__ __forceinline__ void stash_block(float* __restrict__ stx, float* __restrict__ stash_base, int srow, int col, const float (&v)[32], int lane) {
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 8; ++i) *reinterpret_cast<float4*>(stx + lane * SCR_ROW + 4 * i) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int r = 4 * i + (lane >> 3), c4 = 4 * (lane & 7);
        const int sr = __shfl_sync(0xffffffffu, srow, r);
        const float4 x = *reinterpret_cast<const float4*>(stx + r * SCR_ROW + c4);
        if (sr >= 0) __stcs(reinterpret_cast<float4*>(stash_base + (size_t)sr * stash::W + col + c4), x);   // streaming: written once, read once by k_wgrad
    }
}

template <int C, int O>
__device__ void run_decoder(const DecodeParams& P, uint8_t* sm, int dec) {
    using L = Smem<C>;
    constexpr int GCOLS = gcols(C), XCOL = xcol(C), NGC = ngroups(C);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // with a compacted ray list (k_zvals) the tiles walk only the rays that passed the inside filter: no idle rows
    const int total = P.ray_list ? *reinterpret_cast<volatile const int*>(P.ray_count) * P.S : P.P;
    const int ntiles = (total + TM - 1) / TM;
    const uint32_t bar0 = smem_u32(sm + L::BAR), bar_img = bar0 + 16 * NG;
    volatile uint32_t* tmem_ptr = reinterpret_cast<volatile uint32_t*>(sm + L::TMEMPTR);
    volatile int* ticket = reinterpret_cast<volatile int*>(sm + L::TICKET);
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(sm + L::TMEMPTR)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        for (int s = 0; s < NG; ++s) { mbar_init(bar0 + 8 * s, GTHREADS); mbar_init(bar0 + 8 * NG + 8 * s, 1); mbar_init(bar0 + 16 * NG + 8 + 8 * s, GTHREADS); }
        mbar_init(bar_img, 1);
        asm volatile("fence.mbarrier_init.release.cluster;");
        // the decoder's image: one-dimensional bulk TMA copies, completion counted in bytes on bar_img
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_img), "r"((uint32_t)L::IMG) : "memory");
        const uint8_t* src = P.wimg_t5[dec];
        for (int off = 0; off < L::IMG; off += 16384) {
            const uint32_t n = (uint32_t)(L::IMG - off < 16384 ? L::IMG - off : 16384);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_u32(sm + off)), "l"(src + off), "r"(n), "r"(bar_img) : "memory");
        }
    }
    if ((tid & 127) == 0 && (warp >> 2) < NGC) {                    // the first two tiles of every group
        ticket[4 * (warp >> 2)] = (int)atomicAdd(P.tile_ctr + dec, 1ull);
        ticket[4 * (warp >> 2) + 1] = (int)atomicAdd(P.tile_ctr + dec, 1ull);
    }
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = *tmem_ptr;
    constexpr int NO = O == 4 ? 3 : 1;

    if (warp < NG * 4) {
        // ------------------------------------------------------------------ compute threads: one per sample (= TMEM lane)
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS_COMPUTE));
        const int grp = warp >> 2, wq = warp & 3, row = (wq << 5) | lane, q = lane >> 2, t = lane & 3;
        if (grp < NGC) {
        const uint32_t tm = tmem + grp * GCOLS + ((uint32_t)(wq * 32) << 16);
        const uint32_t full = bar0 + 8 * grp, done = bar0 + 8 * NG + 8 * grp;
        const float* bm = reinterpret_cast<const float*>(sm + L::BM);
        const float* wo = reinterpret_cast<const float*>(sm + L::WO);
        const float* woc = reinterpret_cast<const float*>(sm + L::WOC);
        const float* boc = reinterpret_cast<const float*>(sm + L::BOC);
        float* scr = reinterpret_cast<float*>(sm + L::SCR + warp * L::SCRW);
        float* stx = reinterpret_cast<float*>(sm + L::STX + warp * L::STXW);
        // position of this thread's sample in tile `tile` (zero, and inactive, past the end or on a filtered ray)
        auto load_point = [&](int tile, float (&pp)[3], int& s_out) -> bool {
            const int sp = tile * TM + row;
            pp[0] = pp[1] = pp[2] = 0.0f;
            s_out = -1;
            if (tile >= ntiles || sp >= total) return false;
            if (P.pts) { s_out = sp; pp[0] = P.pts[3 * (size_t)sp]; pp[1] = P.pts[3 * (size_t)sp + 1]; pp[2] = P.pts[3 * (size_t)sp + 2]; return true; }
            const int ray = P.ray_list ? P.ray_list[sp / P.S] : sp / P.S;
            const int s = P.ray_list ? ray * P.S + sp % P.S : sp;
            const uint8_t ok = (P.valid && !P.ray_list) ? P.valid[ray] : (uint8_t)1;
            const float z = P.z[s];
            if (!ok) return false;                                       // sin(0 . B) = 0: an inactive row contributes exact zeros
            s_out = s;
#pragma unroll
            for (int a = 0; a < 3; ++a) pp[a] = __fadd_rn(P.rays_o[3 * ray + a], __fmul_rn(P.rays_d[3 * ray + a], z));   // Renderer.cpp:121
            return true;
        };
        // Quad-cooperative gather, pass r of channel chunk cc: the quad q of this warp serves sample 8r + q of the warp and leaves
        // its 8 channels in the scratch row of that sample (a quad fetches a voxel's 128-byte line in one wavefront).
        auto gather_pass = [&](const float (&pp)[3], bool act_lane, int cc, int r) {
            const GridView& G = P.grid[cc == 0 ? dec : 1];              // fine decoder: cat(fine, middle) (MLP.cpp:79-84)
            const int j = 8 * r + q;
            float pj[3];
#pragma unroll
            for (int a = 0; a < 3; ++a) pj[a] = __shfl_sync(0xffffffffu, pp[a], j);
            const bool act = __shfl_sync(0xffffffffu, act_lane ? 1 : 0, j) != 0;
            float c8[8];
            if (act) gather8(G, P.bnd, pj, t, c8);
            else {
#pragma unroll
                for (int i = 0; i < 8; ++i) c8[i] = 0.0f;
            }
            float* dst = scr + cc * 32 * SCR_ROW + j * SCR_ROW + 8 * t;
            *reinterpret_cast<float4*>(dst) = make_float4(c8[0], c8[1], c8[2], c8[3]);
            *reinterpret_cast<float4*>(dst + 4) = make_float4(c8[4], c8[5], c8[6], c8[7]);
        };
        uint32_t step = 0;   // handshakes completed by this group: parity of both barriers
        uint32_t rstep = 0;  // "accumulator read" signals (parity of the ready barrier)
        const uint32_t rdy = bar0 + 16 * NG + 8 + 8 * grp;
        int tile = ticket[4 * grp];
        float p[3];
        int s = -1;
        bool active = load_point(tile, p, s);
        if (tile < ntiles) {   // the first tile's gather runs under the image's TMA copy; later tiles are gathered during the previous tile's layer phase
#pragma unroll
            for (int cc = 0; cc < C / 32; ++cc)
#pragma unroll 1
                for (int r = 0; r < 4; ++r) gather_pass(p, active, cc, r);
        }
        mbar_wait(bar_img, 0);
        for (int round = 0; tile < ntiles; ++round) {
            if (row == 0) ticket[4 * grp + (round + 2) % 3] = (int)atomicAdd(P.tile_ctr + dec, 1ull);   // read after the group barrier of this round
            const int tile_next = ticket[4 * grp + (round + 1) % 3];
            // colour decoder of a training iteration whose weights are optimised: the thread leaves e, c and the relu outputs u_i of its
            // sample in the weight-gradient stash (E | H slots | Cc of params.h; the H slots hold u_i, k_wgrad_finish adds the Fc c + bc part)
            const bool do_stash = O == 4 && P.stash != nullptr;
            const int srow = active ? s : -1;
            // ---- grid features: the owner lane reads its row of the scratch, adds the grid-feature part of the output layer and
            // moves the row to tensor memory as the A operand of every G_i c product
            float outc[NO];
#pragma unroll
            for (int o = 0; o < NO; ++o) outc[o] = boc[o];
            __syncwarp();
#pragma unroll
            for (int cc = 0; cc < C / 32; ++cc) {
                float c[32];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float4 v4 = *reinterpret_cast<const float4*>(scr + cc * 32 * SCR_ROW + lane * SCR_ROW + 4 * i);
                    c[4 * i] = v4.x; c[4 * i + 1] = v4.y; c[4 * i + 2] = v4.z; c[4 * i + 3] = v4.w;
                }
#pragma unroll
                for (int o = 0; o < NO; ++o) {
                    float acc = outc[o];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float4 w4 = *reinterpret_cast<const float4*>(woc + o * C + 32 * cc + 4 * i);
                        acc = fmaf(c[4 * i], w4.x, acc); acc = fmaf(c[4 * i + 1], w4.y, acc); acc = fmaf(c[4 * i + 2], w4.z, acc); acc = fmaf(c[4 * i + 3], w4.w, acc);
                    }
                    outc[o] = acc;
                }
                if (O == 4 && do_stash) stash_block(stx, P.stash, srow, stash::Cc, c, lane);
                store_operand(tm + CCOL + 32 * cc, c);                  // the previous tile's last product has been waited for (layer 4)
            }
            __syncwarp();                                               // the scratch is rewritten by the next tile's gather below
            float pn[3];
            int s_next = -1;
            const bool active_n = load_point(tile_next, pn, s_next);
            // ---- Fourier features, 32 per handshake, this thread's own sample
#pragma unroll 1
            for (int j = 0; j < 3; ++j) {
                float e[32];
#pragma unroll
                for (int k4 = 0; k4 < 8; ++k4) {
                    const int ft = 32 * j + 4 * k4;
                    const float4 b0 = *reinterpret_cast<const float4*>(bm + ft), b1 = *reinterpret_cast<const float4*>(bm + EMBP + ft), b2 = *reinterpret_cast<const float4*>(bm + 2 * EMBP + ft);
                    e[4 * k4] = ff_sin(fmaf(p[2], b2.x, fmaf(p[1], b1.x, p[0] * b0.x)));
                    e[4 * k4 + 1] = ff_sin(fmaf(p[2], b2.y, fmaf(p[1], b1.y, p[0] * b0.y)));
                    e[4 * k4 + 2] = ff_sin(fmaf(p[2], b2.z, fmaf(p[1], b1.z, p[0] * b0.z)));
                    e[4 * k4 + 3] = ff_sin(fmaf(p[2], b2.w, fmaf(p[1], b1.w, p[0] * b0.w)));   // padded columns of B are 0 -> sin(0) = 0
                }
                if (O == 4 && do_stash) stash_block(stx, P.stash, srow, stash::E + 32 * j, e, lane);
                uint32_t w[32];
#pragma unroll
                for (int i = 0; i < 16; ++i) split_f16(e[2 * i], e[2 * i + 1], w[i], w[16 + i]);
                if (j > 0) { mbar_wait(done, (step - 1) & 1); fence_after(); }     // the previous chunk's products have consumed x
                tmem_st32(tm + XCOL, w);
                tmem_st_wait();
                fence_before();
                mbar_arrive(full);
                ++step;
            }
            // ---- five layers: read the accumulator (bias is already in), relu (+ mask), hand u_i back as the next A operand
#pragma unroll 1
            for (int i = 0; i < 5; ++i) {
                mbar_wait(done, (step - 1) & 1);
                fence_after();
                float v[32];
                tmem_ld32(tm + (i == 3 ? ACCS : ACC0), v);
                tmem_ld_wait();
                if (i < 4) {          // the accumulator is in registers: the issuer may start the next layer's grid-feature and bias
                    fence_before();   // products (they do not depend on this layer's result) under the epilogue below
                    mbar_arrive(rdy);
                    ++rstep;
                }
                uint32_t m = 0;
#pragma unroll
                for (int k = 31; k >= 0; --k) {
                    m = __funnelshift_l(__float_as_uint(v[k]), m, 1);      // collects the sign bits: bit k of ~m <-> unit k active
                    v[k] = fmaxf(v[k], 0.0f);
                }
                if (P.masks && active) P.masks[((size_t)(dec - 1) * 5 + i) * P.mask_stride + s] = ~m;
                if (O == 4 && do_stash) stash_block(stx, P.stash, srow, stash::H + HID * i, v, lane);
                if (i < 4) {
                    store_operand(tm + XCOL, v);
                    tmem_st_wait();
                    fence_before();
                    mbar_arrive(full);
                    ++step;
                    if (tile_next < ntiles) {                            // the next tile's gather, pass i, fills the wait for this layer's products
#pragma unroll
                        for (int cc = 0; cc < C / 32; ++cc) gather_pass(pn, active_n, cc, i);
                    }
                } else {
                    float out[NO];
#pragma unroll
                    for (int o = 0; o < NO; ++o) {
                        float acc = outc[o];
#pragma unroll
                        for (int k4 = 0; k4 < 8; ++k4) {
                            const float4 w4 = *reinterpret_cast<const float4*>(wo + o * HID + 4 * k4);
                            acc = fmaf(v[4 * k4], w4.x, acc); acc = fmaf(v[4 * k4 + 1], w4.y, acc); acc = fmaf(v[4 * k4 + 2], w4.z, acc); acc = fmaf(v[4 * k4 + 3], w4.w, acc);
                        }
                        out[o] = acc;
                    }
                    if (active) {
                        if (O == 4) *reinterpret_cast<float4*>(P.out_rgb + 4 * (size_t)s) = make_float4(out[0], out[NO > 1 ? 1 : 0], out[NO > 2 ? 2 : 0], 0.0f);
                        else P.out_occ[dec][s] = out[0];
                    }
                }
            }
            fence_before();       // the accumulator reads above are ordered before the next tile's first arrive
            group_sync(grp);      // row 0's ticket (two rounds ahead) is visible to the group and its issuer
            tile = tile_next; active = active_n; s = s_next;
            p[0] = pn[0]; p[1] = pn[1]; p[2] = pn[2];
        }
        }
    } else {
        // ------------------------------------------------------------------ MMA issuer of one tile group (lane 0 issues)
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_ISSUE));
        const int grp = warp - NG * 4;
        if (grp < NGC) {
        const uint32_t tm = tmem + grp * GCOLS;
        const uint32_t full = bar0 + 8 * grp, done = bar0 + 8 * NG + 8 * grp;
        constexpr uint32_t I32 = make_idesc(128, 32), I64 = make_idesc(128, 64);
        constexpr int NC = C / 32;
        // every shared-memory descriptor is the image's base descriptor plus a byte offset >> 4 (few live registers: the issuer
        // runs on 32 registers per thread)
        const uint64_t dbase = make_desc(smem_u32(sm));
        uint32_t step = 0, rstep = 0;
        const uint32_t rdy = bar0 + 16 * NG + 8 + 8 * grp;
        mbar_wait(bar_img, 0);
        for (int round = 0, tile = ticket[4 * grp]; tile < ntiles; ++round) {
            if (lane == 0) {
#pragma unroll 1
                for (int j = 0; j < 3; ++j) {           // [acc0 | accS] (+)= e_j [W0 ; W3E]_j^T
                    mbar_wait(full, step & 1); fence_after();
                    issue3(tm + ACC0, tm + XCOL, dbase + (uint64_t)((L::WE + j * 8192) >> 4), I64, j == 0);
                    if (j == 0) mma_ss(tm + ACC0, dbase + (L::ONE >> 4), dbase + (L::BT >> 4), I64, 1u);                 // + b'_0 | b'_3
                    mma_commit(done); ++step;
                }
#pragma unroll 1
                for (int l = 0; l < 4; ++l) {           // layer l+1: acc = c G_l^T + b' (early: under the epilogue of layer l) + u_l W_{l+1}^T (late); layer 3 accumulates into accS
                    const uint32_t d = tm + (l == 2 ? ACCS : ACC0);
                    mbar_wait(rdy, rstep & 1); fence_after(); ++rstep;      // every thread has the previous accumulator in registers
#pragma unroll 1
                    for (int cc = 0; cc < NC; ++cc) issue3(d, tm + CCOL + 32 * cc, dbase + (uint64_t)((L::G + (l * NC + cc) * 4096) >> 4), I32, l != 2 && cc == 0);
                    if (l != 2) mma_ss(d, dbase + (L::ONE >> 4), dbase + (uint64_t)((L::BT >> 4) + 2 * (l == 3 ? 3 : l + 1)), I32, 1u);   // + b'_{l+1}
                    mbar_wait(full, step & 1); fence_after();
                    issue3(d, tm + XCOL, dbase + (uint64_t)((L::WH + l * 4096) >> 4), I32, false);
                    mma_commit(done); ++step;
                }
            } else { step += 7; rstep += 4; }
            __syncwarp();
            group_sync(grp);
            tile = ticket[4 * grp + (round + 1) % 3];
        }
        }
    }
    fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u));
    if (tid == 0 && P.ray_list) {      // every CTA has read the list length: the last one to finish clears it for the next k_zvals
        __threadfence();
        if (atomicAdd(P.ray_count + 1, 1) == (int)gridDim.x - 1) { P.ray_count[0] = 0; P.ray_count[1] = 0; }
    }
}

__global__ void __launch_bounds__(THREADS, 1) k_decode_fwd_t5(const DecodeParams P) {
    extern __shared__ __align__(1024) uint8_t sm[];                 // used in place (shared state space known to the compiler: LDS, not generic loads)
    if (smem_u32(sm) & 1023u) __trap();                             // SWIZZLE_128B tiles need 1024-byte alignment
    int dec = 1;
#pragma unroll
    for (int d = 2; d < 4; ++d) if ((int)blockIdx.x >= P.cta_begin[d]) dec = d;
    if (dec == 1) run_decoder<32, 1>(P, sm, 1);
    else if (dec == 2) run_decoder<64, 1>(P, sm, 2);
    else run_decoder<32, 4>(P, sm, 3);
}

// Builds the shared-memory image of decoder d (blockIdx.y = d - 1) in global memory; run after k_compose whenever the weights change.
struct T5ImgParams { const float* flat[4]; const float* comp[4]; uint8_t* img[4]; int mask; };
__global__ void __launch_bounds__(512) k_build_t5img(T5ImgParams W) {
    const int d = 1 + blockIdx.y;
    if (!((W.mask >> d) & 1)) return;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nthr = gridDim.x * blockDim.x;
    if (d == 1) stage<32, 1>(W.img[1], W.flat[1], W.comp[1], tid, nthr);
    else if (d == 2) stage<64, 1>(W.img[2], W.flat[2], W.comp[2], tid, nthr);
    else stage<32, 4>(W.img[3], W.flat[3], W.comp[3], tid, nthr);
}

}  // namespace t5

size_t t5_img_bytes(int which) { return which == 2 ? t5::Smem<64>::IMG : t5::Smem<32>::IMG; }
cudaError_t launch_build_t5img(const float* const flat[4], const float* const comp[4], uint8_t* const img[4], int mask, cudaStream_t st) {
    t5::T5ImgParams W;
    for (int d = 0; d < 4; ++d) { W.flat[d] = flat[d]; W.comp[d] = comp[d]; W.img[d] = img[d]; }
    W.mask = mask;
    t5::k_build_t5img<<<dim3(8, 3), 512, 0, st>>>(W);
    return cudaGetLastError();
}

cudaError_t launch_decode_fwd_t5(const DecodeParams& P, int grid, cudaStream_t st) {
    const size_t smem = (size_t)(t5::Smem<64>::TOTAL > t5::Smem<32>::TOTAL ? t5::Smem<64>::TOTAL : t5::Smem<32>::TOTAL);
    static unsigned attr_done = 0;      // per device: function attributes are per-device state
    int dev = 0; cudaGetDevice(&dev);
    if (!((attr_done >> (dev & 31)) & 1u)) {
        cudaError_t e = cudaFuncSetAttribute(t5::k_decode_fwd_t5, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attr_done |= 1u << (dev & 31);
    }
    t5::k_decode_fwd_t5<<<grid, t5::THREADS, smem, st>>>(P);
    return cudaGetLastError();
}

}  // namespace nsb
