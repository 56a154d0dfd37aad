// Split-K weight-gradient kernel for the colour decoder (the only decoder the reference optimises by default:
// fix_fine = True, fix_color = False, Mapper.cpp:292-301).  The backward decoder kernel stashes, per sample, the
// layer inputs X and the layer-output gradients G; every parameter gradient is then a tall-skinny product
//     dW[a][b] = sum_s L[s][a] * R[s][b]
// with the sample index as the contraction dimension.  Each CTA walks a contiguous range of samples; each warp owns one
// task = up to four 16-row A tiles x ONE 32-column B block (accumulators in registers), one warp sums the bias columns;
// the per-CTA partial sums are added to the flat gradient with fp32 reductions at the end.  The 16 stash rows of a k-step are one
// contiguous 45.6 KB block: a producer lane moves them with ONE bulk TMA copy per stage (cp.async.bulk -> mbarrier complete_tx) and
// the consumer warps run decoupled behind full / empty mbarriers -- no CTA-wide barrier per k-step (it cost 2.8 stalled warps per
// issue slot: the warps' tasks differ by 4x in size), no per-thread copy instructions.
#include "decode.cuh"
#include "params.h"
#include <cstring>
#include <initializer_list>

#ifndef NSB_MBAR_HINT_NS
#define NSB_MBAR_HINT_NS 4000   // suspend-time hint of mbarrier.try_wait in ns (0: the loop spins; measured neutral in time, fewer issued instructions)
#endif

namespace nsb {

constexpr int WG_WARPS = 16;
constexpr int WG_MAXL = 4;          // A tiles (16 stash columns each) a warp multiplies with its one B block (32 stash columns)
constexpr int WG_SCRATCH = 1 << 24;  // dst tag of the tasks that accumulate into the M scratch ([5][32][32] floats: M_1..M_4, M_out)
constexpr int WG_NBIAS = 6;         // column-sum blocks of the bias warp (b_0..b_4, bo)

struct WGroup {
    int L;             // stash column of row 0 of the A tile
    int dst, ld;       // flat-gradient offset of element (a_lo, 0) and its row stride; dst >= WG_SCRATCH: offset into the M scratch instead
                       // (M_i = sum_s g_u_i (x) c, finished by k_wgrad_finish)
    int a_lo, a_hi;    // valid rows of the A tile (relative to L)
};
// One task per warp: dW[L rows][R cols] += sum_s stash[s][L + a] * stash[s][R + b] for up to four A tiles that share ONE B
// block, so that the B operand is split into fp16 hi/lo once per k-step and reused (the split, not the MMA, dominates the
// instruction count).  kind 1 is the bias warp: plain fp32 column sums of six blocks (b_i = sum GU_i, bo = sum GO; bc_i follows in k_wgrad_finish).
struct WTask {
    int kind;          // 0 idle, 1 bias sums, 2 MMA task
    int nL, R, nb;     // number of A tiles, stash column of the B block, valid B columns (<= 32)
    WGroup g[WG_MAXL];
    int bias_col[WG_NBIAS], bias_dst[WG_NBIAS], bias_n[WG_NBIAS];
};
struct WGTable { WTask t[WG_WARPS]; };

static WGTable build_table() {
    WGTable T; memset(&T, 0, sizeof T);
    const DecFlat f = DecFlat::make(32, 4);
    auto mma = [&](int warp, int R, int nb, std::initializer_list<WGroup> gs) {
        WTask& t = T.t[warp]; t.kind = 2; t.R = R; t.nb = nb; t.nL = 0;
        for (const WGroup& g : gs) t.g[t.nL++] = g;
    };
    auto G = [&](int L, int dst, int ld, int a_hi = 16, int a_lo = 0) { return WGroup{L, dst, ld, a_lo, a_hi}; };
    const int GU = stash::GU, LD3 = EMB + HID, MS = WG_SCRATCH;
    // Warps are spread over the four schedulers (warp & 3) by their HMMA count per k-step (48 / 24 / 12 per task, 396 in all): 96 + 96 +
    // 108 + 96 (measured: rebalancing from 48 / 120 / 120 / 108 changed nothing -- the ring is paced by per-warp latency chains, not by
    // one scheduler's HMMA count).
    // scheduler 0: E0 x {GU0, GU3} (48), H0 x GU1 (W1, 24), Cc x GO (M_out, 12), H4 x GO (Wo, 12)
    mma(0, stash::E + 0, 32, {G(GU + 0, f.W[0], EMB), G(GU + 16, f.W[0] + 16 * EMB, EMB), G(GU + 96, f.W[3], LD3), G(GU + 112, f.W[3] + 16 * LD3, LD3)});
    mma(4, stash::H + 0, 32, {G(GU + 32, f.W[1], 32), G(GU + 48, f.W[1] + 16 * 32, 32)});
    mma(8, stash::Cc, 32, {G(stash::GO, MS + 4 * 1024, 32, 4)});
    mma(12, stash::H + 128, 32, {G(stash::GO, f.Wo, 32, 4)});
    // scheduler 1: E1, E2 x {GU0, GU3} (48 + 48), the bias sums (no MMAs), and the idle warp 13 = the TMA producer
    mma(1, stash::E + 32, 32, {G(GU + 0, f.W[0] + 32, EMB), G(GU + 16, f.W[0] + 16 * EMB + 32, EMB), G(GU + 96, f.W[3] + 32, LD3), G(GU + 112, f.W[3] + 16 * LD3 + 32, LD3)});
    mma(5, stash::E + 64, EMB - 64, {G(GU + 0, f.W[0] + 64, EMB), G(GU + 16, f.W[0] + 16 * EMB + 64, EMB),
                                     G(GU + 96, f.W[3] + 64, LD3), G(GU + 112, f.W[3] + 16 * LD3 + 64, LD3)});
    {
        WTask& t = T.t[9]; t.kind = 1;
        for (int i = 0; i < 5; ++i) { t.bias_col[i] = GU + 32 * i; t.bias_dst[i] = f.b[i]; t.bias_n[i] = 32; }
        t.bias_col[5] = stash::GO; t.bias_dst[5] = f.bo; t.bias_n[5] = 4;
    }
    // scheduler 2: Cc x {GU1, GU2} (M_1, M_2, 48), H1 x GU2 (W2, 24), H2 x GU3 (W3 hidden columns, 24), GE0 x p (dB, 12)
    mma(2, stash::Cc, 32, {G(GU + 32, MS + 0, 32), G(GU + 48, MS + 16 * 32, 32), G(GU + 64, MS + 1024, 32), G(GU + 80, MS + 1024 + 16 * 32, 32)});
    mma(6, stash::H + 32, 32, {G(GU + 64, f.W[2], 32), G(GU + 80, f.W[2] + 16 * 32, 32)});
    mma(10, stash::H + 64, 32, {G(GU + 96, f.W[3] + EMB, LD3), G(GU + 112, f.W[3] + 16 * LD3 + EMB, LD3)});
    mma(14, stash::GE + 0, 32, {G(stash::Pp, f.B, EMB, 3)});
    // scheduler 3: Cc x {GU3, GU4} (M_3, M_4, 48), H3 x GU4 (W4, 24), GE1, GE2 x p (12 + 12)
    mma(3, stash::Cc, 32, {G(GU + 96, MS + 2048, 32), G(GU + 112, MS + 2048 + 16 * 32, 32), G(GU + 128, MS + 3072, 32), G(GU + 144, MS + 3072 + 16 * 32, 32)});
    mma(7, stash::H + 96, 32, {G(GU + 128, f.W[4], 32), G(GU + 144, f.W[4] + 16 * 32, 32)});
    mma(11, stash::GE + 32, 32, {G(stash::Pp, f.B + 32, EMB, 3)});
    mma(15, stash::GE + 64, EMB - 64, {G(stash::Pp, f.B + 64, EMB, 3)});
    return T;
}

__constant__ WGTable c_wg;

constexpr int WG_STAGES = 6;                       // ring depth (k-steps in flight): 6 x 35.5 KB
constexpr int WG_KROWS = 16;                       // samples per k-step (MMA k = 16)
constexpr int WG_STAGE_FLOATS = WG_KROWS * stash::W + 32;   // +32: the last B block of a row may read past it (values unused)

__device__ __forceinline__ uint32_t wg_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void wg_mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void wg_mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void wg_mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred P1;\n\tWAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n\t"      // suspend-time hint: sleep in hardware instead of spinning
        "@P1 bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(bar), "r"(parity), "r"((unsigned)NSB_MBAR_HINT_NS) : "memory");
}

// The CTA streams its sample range through a shared-memory ring (one stage = the 16 stash rows of a k-step, 45.6 KB,
// fetched once with 16-byte cp.async), so every stash element crosses L2/HBM exactly once and the MMA fragments come from
// conflict-free LDS.64 / LDS.128 (row stride 712 = 8 mod 32 floats).  k slots (2t, 2t+1, 2t+8, 2t+9) of lane t <-> ring
// rows (t, t+4, t+8, t+12): any bijection works as long as both operands use it.  A rows g / g+8 <-> stash columns L + 2g,
// L + 2g + 1; B column g of n-tile j <-> stash column R + 4g + j (one float4 per row serves the four n-tiles).
template <bool P3>
__global__ void __launch_bounds__(WG_WARPS * 32) k_wgrad(const float* __restrict__ st, const uint8_t* __restrict__ valid,
                                                         int P, int S, float* __restrict__ dflat, float* __restrict__ mscr) {
    extern __shared__ __align__(128) float ring[];
    __shared__ __align__(8) unsigned long long bars[2 * WG_STAGES];       // full[stage], empty[stage]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int kind = c_wg.t[warp].kind, nL = c_wg.t[warp].nL;
    float acc[WG_MAXL][4][4];     // bias warp: acc[q][j][e] doubles as 44 column-sum accumulators
#pragma unroll
    for (int q = 0; q < WG_MAXL; ++q)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[q][j][0] = acc[q][j][1] = acc[q][j][2] = acc[q][j][3] = 0.0f;
    int Lc[WG_MAXL];
#pragma unroll
    for (int q = 0; q < WG_MAXL; ++q) Lc[q] = c_wg.t[warp].g[q < nL ? q : 0].L + 2 * g;
    const int Rc = c_wg.t[warp].R + 4 * g;
    const int nks = P / WG_KROWS;
    const int per = (nks + gridDim.x - 1) / gridDim.x;
    const int k_lo = blockIdx.x * per, k_hi = min(nks, k_lo + per);
    const int n_my = max(0, k_hi - k_lo);
    const uint32_t bar0 = wg_smem_u32(bars);
    if (threadIdx.x == 0) {
        int consumers = 0;
        for (int w = 0; w < WG_WARPS; ++w) consumers += c_wg.t[w].kind != 0;
        for (int s_ = 0; s_ < WG_STAGES; ++s_) { wg_mbar_init(bar0 + 8 * s_, 1); wg_mbar_init(bar0 + 8 * (WG_STAGES + s_), consumers); }
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    __syncthreads();
    constexpr uint32_t STAGE_BYTES = WG_KROWS * stash::W * 4;
    if (kind == 0) {
        // ---- producer: the first idle warp's lane 0 streams the CTA's sample range, one bulk TMA copy per k-step
        bool first_idle = true;
        for (int w = 0; w < warp; ++w) if (c_wg.t[w].kind == 0) first_idle = false;
        if (first_idle && lane == 0) {
            // the stash is read exactly once: evict-first in L2, so that the grids and the arenas of the optimiser step stay resident
            uint64_t pol;
            asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
            for (int i = 0; i < n_my; ++i) {
                const int slot = i % WG_STAGES;
                if (i >= WG_STAGES) wg_mbar_wait(bar0 + 8 * (WG_STAGES + slot), ((i / WG_STAGES) - 1) & 1);    // every consumer warp is done with the slot
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar0 + 8 * slot), "r"(STAGE_BYTES) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                             ::"r"(wg_smem_u32(ring + slot * WG_STAGE_FLOATS)), "l"(st + (size_t)(k_lo + i) * WG_KROWS * stash::W), "r"(STAGE_BYTES), "r"(bar0 + 8 * slot), "l"(pol) : "memory");
            }
        }
        return;
    }
    for (int i = 0; i < n_my; ++i) {
        const int slot = i % WG_STAGES;
        const int s0 = (k_lo + i) * WG_KROWS;
        const bool live = !(valid && !valid[s0 / S]);   // rows of rays dropped by the inside filter were never written
        wg_mbar_wait(bar0 + 8 * slot, (i / WG_STAGES) & 1);   // the k-step's 16 rows have landed
        const float* r0 = ring + slot * WG_STAGE_FLOATS + t * stash::W;
        if (live && kind == 2) {
            float4 bv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) bv[u] = *reinterpret_cast<const float4*>(r0 + 4 * u * stash::W + Rc);
            uint32_t bh[4][2], bl[4][2];               // the B block, split once for all A tiles of the task
            if (P3) {
                split_f16(bv[0].x, bv[1].x, bh[0][0], bl[0][0]); split_f16(bv[2].x, bv[3].x, bh[0][1], bl[0][1]);
                split_f16(bv[0].y, bv[1].y, bh[1][0], bl[1][0]); split_f16(bv[2].y, bv[3].y, bh[1][1], bl[1][1]);
                split_f16(bv[0].z, bv[1].z, bh[2][0], bl[2][0]); split_f16(bv[2].z, bv[3].z, bh[2][1], bl[2][1]);
                split_f16(bv[0].w, bv[1].w, bh[3][0], bl[3][0]); split_f16(bv[2].w, bv[3].w, bh[3][1], bl[3][1]);
            } else {
                bh[0][0] = pack_f16(bv[0].x, bv[1].x); bh[0][1] = pack_f16(bv[2].x, bv[3].x);
                bh[1][0] = pack_f16(bv[0].y, bv[1].y); bh[1][1] = pack_f16(bv[2].y, bv[3].y);
                bh[2][0] = pack_f16(bv[0].z, bv[1].z); bh[2][1] = pack_f16(bv[2].z, bv[3].z);
                bh[3][0] = pack_f16(bv[0].w, bv[1].w); bh[3][1] = pack_f16(bv[2].w, bv[3].w);
            }
#pragma unroll
            for (int q = 0; q < WG_MAXL; ++q) {
                if (q >= nL) break;
                float2 av[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) av[u] = *reinterpret_cast<const float2*>(r0 + 4 * u * stash::W + Lc[q]);
                AFrag<P3> a;
                a.set(av[0].x, av[1].x, av[2].x, av[3].x, av[0].y, av[1].y, av[2].y, av[3].y);
                // product by product over the four n-tiles: consecutive HMMAs hit different accumulators (the three products of one
                // accumulator are a dependent chain)
                if (P3) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) mma_f16(acc[q][j], a.lo, bh[j][0], bh[j][1]);
#pragma unroll
                    for (int j = 0; j < 4; ++j) mma_f16(acc[q][j], a.hi, bl[j][0], bl[j][1]);
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) mma_f16(acc[q][j], a.hi, bh[j][0], bh[j][1]);
            }
        } else if (live && kind == 1) {
            float* sums = &acc[0][0][0];
#pragma unroll
            for (int b = 0; b < WG_NBIAS; ++b) {
                const int col = c_wg.t[warp].bias_col[b] + 4 * g;
                if (4 * g < c_wg.t[warp].bias_n[b]) {
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const float4 v = *reinterpret_cast<const float4*>(r0 + 4 * u * stash::W + col);
                        sums[4 * b] += v.x; sums[4 * b + 1] += v.y; sums[4 * b + 2] += v.z; sums[4 * b + 3] += v.w;
                    }
                }
            }
        }
        __syncwarp();                                  // every lane's reads of the slot are done
        if (lane == 0) wg_mbar_arrive(bar0 + 8 * (WG_STAGES + slot));
    }
    if (kind == 2) {
        const int nb = c_wg.t[warp].nb;
#pragma unroll
        for (int q = 0; q < WG_MAXL; ++q) {
            if (q >= nL) break;
            const WGroup G = c_wg.t[warp].g[q];
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int a = 2 * g + (e >> 1), b = 4 * (2 * t + (e & 1)) + j;
                    if (a >= G.a_lo && a < G.a_hi && b < nb) atomicAdd((G.dst >= WG_SCRATCH ? mscr + (G.dst - WG_SCRATCH) : dflat + G.dst) + (a - G.a_lo) * G.ld + b, acc[q][j][e]);
                }
        }
    } else if (kind == 1) {
        float* sums = &acc[0][0][0];
#pragma unroll
        for (int b = 0; b < WG_NBIAS; ++b) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float v = quad_sum(sums[4 * b + k]);          // the four lanes of a quad hold different sample rows
                if (t == 0 && 4 * g + k < c_wg.t[warp].bias_n[b]) atomicAdd(dflat + c_wg.t[warp].bias_dst[b] + 4 * g + k, v);
            }
        }
    }
}

// d Fc_{i-1} = W_i(hidden columns)^T M_i and d bc_{i-1} = W_i^T d b_i for i = 1..4, d Fc_4 = Wo^T M_out and d bc_4 = Wo^T d bo: the gradients at the
// block outputs are linear images of the g_u sums k_wgrad has just produced.  One block per Fc matrix; clears the M scratch for the next iteration.
__global__ void __launch_bounds__(1024) k_wgrad_finish(const float* __restrict__ flat, float* __restrict__ dflat, float* __restrict__ mscr, int h_is_u) {
    const DecFlat f = DecFlat::make(32, 4);
    const int i = blockIdx.x;                       // Fc_i, i = 0..4
    __shared__ float sW[32][33], sM[32][33], sb[32];
    const int no = i == 4 ? 3 : 32;                 // contraction length (the colour decoder's 4th output is overwritten: no gradient)
    for (int k = threadIdx.x; k < 1024; k += blockDim.x) {
        const int o = k / 32, m = k % 32;
        float w = 0.0f, mm = 0.0f;
        if (o < no) {
            w = i == 4 ? flat[f.Wo + o * HID + m] : i == 2 ? flat[f.W[3] + o * (EMB + HID) + EMB + m] : flat[f.W[i + 1] + o * HID + m];
            mm = mscr[i * 1024 + o * 32 + m];       // here m is the channel index of M
        }
        sW[o][m] = w; sM[o][m] = mm;
    }
    if (threadIdx.x < 32) sb[threadIdx.x] = threadIdx.x < no ? (i == 4 ? dflat[f.bo + threadIdx.x] : dflat[f.b[i + 1] + threadIdx.x]) : 0.0f;
    __syncthreads();
    for (int k = threadIdx.x; k < 1024; k += blockDim.x) {
        const int m = k / 32, ch = k % 32;
        float acc = 0.0f;
#pragma unroll 8
        for (int o = 0; o < 32; ++o) acc = fmaf(sW[o][m], sM[o][ch], acc);
        dflat[f.Fc[i] + m * 32 + ch] += acc;
    }
    if (threadIdx.x < 32) {
        float acc = 0.0f;
        for (int o = 0; o < 32; ++o) acc = fmaf(sW[o][threadIdx.x], sb[o], acc);
        dflat[f.bc[i] + threadIdx.x] += acc;
    }
    if (h_is_u) {
        // The tcgen05 forward stashes the relu outputs u_i in the H slots, not the block outputs h_{i+1} = u_i + Fc_i c + bc_i, so k_wgrad's
        // H x g_u products miss  (sum_s g_u (x) c) Fc_i^T + (sum_s g_u) bc_i^T = M Fc_i^T + db bc_i^T  -- added here, same M and db as above.
        __shared__ float sF[32][33], sbc[32];
        for (int k = threadIdx.x; k < 1024; k += blockDim.x) sF[k / 32][k % 32] = flat[f.Fc[i] + k];        // Fc_i[m][ch]
        if (threadIdx.x < 32) sbc[threadIdx.x] = flat[f.bc[i] + threadIdx.x];
        __syncthreads();
        for (int k = threadIdx.x; k < 1024; k += blockDim.x) {
            const int o = k / 32, m = k % 32;
            if (o >= no) continue;
            float acc = sb[o] * sbc[m];
#pragma unroll 8
            for (int ch = 0; ch < 32; ++ch) acc = fmaf(sM[o][ch], sF[m][ch], acc);
            float* dst = i == 4 ? dflat + f.Wo + o * HID + m : i == 2 ? dflat + f.W[3] + o * (EMB + HID) + EMB + m : dflat + f.W[i + 1] + o * HID + m;
            *dst += acc;
        }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < 1024; k += blockDim.x) mscr[i * 1024 + k] = 0.0f;
}

// Per-device one-time setup (task table in constant memory, shared-memory attribute); called from nsb_create so that the launches
// themselves are plain kernel launches (capturable into a CUDA graph).
cudaError_t wgrad_init() {
    static unsigned init = 0;           // per device: the task table lives in that device's constant memory
    int dev = 0; cudaGetDevice(&dev);
    const size_t smem = sizeof(float) * WG_STAGES * WG_STAGE_FLOATS;
    if (!((init >> (dev & 31)) & 1u)) {
        const WGTable T = build_table();
        cudaError_t e = cudaMemcpyToSymbol(c_wg, &T, sizeof(T));
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_wgrad<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_wgrad<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        init |= 1u << (dev & 31);
    }
    return cudaSuccess;
}

// flat: the colour decoder's parameters (W^T of the finishing step); mscr: [5][32][32] floats, zero on entry and on exit; h_is_u: the stash's
// H slots hold the relu outputs (tcgen05 forward) instead of the block outputs (warp-MMA forward)
cudaError_t launch_wgrad(const float* stash_buf, const uint8_t* valid, int P, int S, const float* flat, float* dflat, float* mscr, int h_is_u, int precision, int grid, cudaStream_t st) {
    const size_t smem = sizeof(float) * WG_STAGES * WG_STAGE_FLOATS;
    { const cudaError_t e = wgrad_init(); if (e != cudaSuccess) return e; }
    if (precision == 0) k_wgrad<true><<<grid, WG_WARPS * 32, smem, st>>>(stash_buf, valid, P, S, dflat, mscr);
    else k_wgrad<false><<<grid, WG_WARPS * 32, smem, st>>>(stash_buf, valid, P, S, dflat, mscr);
    k_wgrad_finish<<<5, 1024, 0, st>>>(flat, dflat, mscr, h_is_u);
    return cudaGetLastError();
}

}  // namespace nsb
