// Forward decoder kernel on the 5th-generation tensor cores, fp16 two-way split (tcgen05.mma kind::f16, accumulators in
// tensor memory).  Second generation of decode_fwd_tc.cu (kind::tf32): the same fp32-grade product as the warp-MMA path
//     D += a_lo.b_hi + a_hi.b_lo + a_hi.b_hi,   x = hi + lo,  hi = fp16(x), lo = fp16(x - hi)
// with every operand row stored as ONE 128-byte SWIZZLE_128B row  [ hi (32 halves) | lo (32 halves) ]  so that the three
// partial products are three descriptor offsets (A+64B x B, A x B+64B, A x B) and a 32-wide layer is six UMMA instructions.
//
// One CTA per SM keeps ONE decoder resident (weights 56-80 KB) and runs THREE 128-sample tiles concurrently:
//   warps 4s..4s+3  (tile slot s)   one thread per sample = one TMEM lane: quad-cooperative trilinear gather (LDG.256, a quad
//                                   fetches a voxel line in one wavefront), Fourier features, per-layer epilogues (tcgen05.ld,
//                                   bias, relu, mask, fp16 split, swizzled 16-byte stores into the slot's A tile)
//   warps 12..14                    one elected lane per slot issues the tcgen05.mma stream and commits to an mbarrier
// Algebra (k_compose, shared with decode_fwd_tc.cu): a_{i+1} = W_{i+1} relu(a_i) + G_i c + b'_{i+1}, G_i = W_{i+1} Fc_i.  All four
// grid-feature terms G_i c are computed FIRST, as one N = 160 product straight into the accumulator columns of layers 1..4 (the
// first 32 rows of that B tile are zeros, which also clears layer 0's accumulator), so the A operand of every later step is just
// the embedding chunk or u_i = relu(a_i), and each layer costs one accumulator read.
// Replaces NICE::forward / MLP::forward (NICE.cpp:16-51, MLP.cpp:76-102); selected with NSB_TCGEN05=2 when no wgrad stash is needed.
#include "decode.cuh"
#include "params.h"

namespace nsb {
namespace tc16 {

constexpr int TM = 128;                        // samples per tile (UMMA M)
constexpr int TILES = 3;                       // tile slots per CTA (160 TMEM columns each)
constexpr int TTHREADS = 128;                  // compute threads per slot: one per sample
constexpr int CTHREADS = TILES * TTHREADS;
constexpr int THREADS = CTHREADS + 32 * TILES; // + one issuer warp per slot
constexpr int SLOT_COLS = 160;
// accumulator columns inside a slot: layer 0 and the skip layer are adjacent (one N = 64 embedding product), then 1, 2, 4
__host__ __device__ constexpr int acc_col(int layer) { return layer == 0 ? 0 : layer == 3 ? 32 : layer == 1 ? 64 : layer == 2 ? 96 : 128; }

// composed weights (global, per decoder), layout of k_compose: G[4][32][C] | bp[5][32] | woc[4][C] | boc[4]
__host__ __device__ constexpr int comp_G(int) { return 0; }
__host__ __device__ constexpr int comp_bp(int C) { return 4 * HID * C; }
__host__ __device__ constexpr int comp_woc(int C) { return comp_bp(C) + 5 * HID; }
__host__ __device__ constexpr int comp_boc(int C) { return comp_woc(C) + 4 * C; }

template <int C>
struct Smem {   // bytes; every UMMA tile starts on a multiple of 1024 B, rows are 128 B
    static constexpr int WE = 0;                                   // 3 chunks x [64 rows]: rows 0-31 W0, 32-63 W3 (embedding columns)
    static constexpr int WH = WE + 3 * 64 * 128;                   // 4 x [32 rows]: W1, W2, W3 (hidden columns), W4
    static constexpr int GC = WH + 4 * 32 * 128;                   // C/32 chunks x [160 rows]: zeros | G_2 | G_0 | G_1 | G_3
    static constexpr int A = GC + (C / 32) * 160 * 128;            // TILES x [128 rows]
    static constexpr int BM = A + TILES * TM * 128;                // Fourier matrix [3][96] fp32
    static constexpr int BP = BM + 3 * EMBP * 4;                   // b'[5][32]
    static constexpr int WO = BP + 5 * HID * 4;                    // Wo[4][32]
    static constexpr int WOC = WO + 4 * HID * 4;                   // (Wo Fc_4)[4][C]
    static constexpr int BOC = WOC + 4 * C * 4;                    // const[4]
    static constexpr int XCH = BOC + 16;                           // TILES x 128 x float4: grid-feature part of the output layer
    static constexpr int BAR = XCH + TILES * TM * 16;              // mbarriers: full_A[TILES], mma_done[TILES]
    static constexpr int TMEMPTR = BAR + 2 * TILES * 8;
    static constexpr int TOTAL = TMEMPTR + 16;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred P1;\n\tWAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {   // K-major SWIZZLE_128B, SBO 1024 B, version 1
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {   // kind::f16: A, B = F16 (format 0), D = F32, both K-major
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]),
                   "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]),
                   "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// byte offset of the 16-byte chunk `c` (0..3 hi, 4..7 lo) of row r inside a SWIZZLE_128B tile
__device__ __forceinline__ int chunk_off(int r, int c) { return (r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4); }

// 8 fp32 values -> one 16-byte chunk of fp16 hi parts and one of lo parts
__device__ __forceinline__ void split8(const float* v, uint4& hi, uint4& lo) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) split_f16(v[2 * i], v[2 * i + 1], h[i], l[i]);
    hi = make_uint4(h[0], h[1], h[2], h[3]); lo = make_uint4(l[0], l[1], l[2], l[3]);
}
// columns [8k, 8k+8) of a thread's 32-wide row -> A tile (hi chunk k, lo chunk 4 + k)
__device__ __forceinline__ void store_row32(uint8_t* a_tile, int row, const float (&v)[32]) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        uint4 hi, lo;
        split8(v + 8 * k, hi, lo);
        *reinterpret_cast<uint4*>(a_tile + chunk_off(row, k)) = hi;
        *reinterpret_cast<uint4*>(a_tile + chunk_off(row, 4 + k)) = lo;
    }
}

// one element of a weight tile: row r, logical input index k (0..31) -> hi at half k, lo at half 32 + k
__device__ __forceinline__ void put_w(uint8_t* tile, int r, int k, float v) {
    const __half h = __float2half_rn(v);
    const __half l = __float2half_rn(v - __half2float(h));
    *reinterpret_cast<__half*>(tile + chunk_off(r, k >> 3) + (k & 7) * 2) = h;
    *reinterpret_cast<__half*>(tile + chunk_off(r, 4 + (k >> 3)) + (k & 7) * 2) = l;
}

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
    for (int idx = tid; idx < (C / 32) * 160 * 32; idx += nthr) {
        const int cc = idx / (160 * 32), r = (idx / 32) % 160, k = idx % 32;
        float v = 0.0f;
        if (r >= 32) {
            const int blk = (r - 32) / 32, o = (r - 32) % 32;
            const int i = blk == 0 ? 2 : blk == 1 ? 0 : blk == 2 ? 1 : 3;      // rows: zeros | G_2 | G_0 | G_1 | G_3
            v = comp[comp_G(C) + (i * HID + o) * C + 32 * cc + k];
        }
        put_w(sm + L::GC + cc * 20480, r, k, v);
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

// fp32-grade product over one 32-input block: D += A_lo B_hi + A_hi B_lo + A_hi B_hi  (descriptor units of 16 B: +2 = one k16
// step, +4 = the lo half of the 128-byte row)
__device__ __forceinline__ void issue3(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, bool zero_first) {
#pragma unroll
    for (int k = 0; k < 2; ++k) mma_ss(d, a + 4 + 2 * k, b + 2 * k, idesc, (zero_first && k == 0) ? 0u : 1u);
#pragma unroll
    for (int k = 0; k < 2; ++k) mma_ss(d, a + 2 * k, b + 4 + 2 * k, idesc, 1u);
#pragma unroll
    for (int k = 0; k < 2; ++k) mma_ss(d, a + 2 * k, b + 2 * k, idesc, 1u);
}

template <int C, int O>
__device__ void run_decoder(const DecodeParams& P, uint8_t* sm, int dec, int cta, int ncta) {
    using L = Smem<C>;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ntiles = (P.P + TM - 1) / TM;
    const uint32_t bar0 = smem_u32(sm + L::BAR);
    volatile uint32_t* tmem_ptr = reinterpret_cast<volatile uint32_t*>(sm + L::TMEMPTR);
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(sm + L::TMEMPTR)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        for (int s = 0; s < TILES; ++s) { mbar_init(bar0 + 8 * s, TTHREADS); mbar_init(bar0 + 8 * TILES + 8 * s, 1); }
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    stage<C, O>(sm, P.dec_flat[dec], P.comp[dec], tid, THREADS);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = *tmem_ptr;
    constexpr int NO = O == 4 ? 3 : 1;

    if (warp < TILES * 4) {
        // ------------------------------------------------------------------ compute threads: one per sample (= TMEM lane)
        const int slot = warp >> 2, wq = warp & 3, row = (wq << 5) | lane, q = lane >> 2, t = lane & 3;
        const uint32_t tm = tmem + slot * SLOT_COLS + ((uint32_t)(wq * 32) << 16);
        const uint32_t full_a = bar0 + 8 * slot, mma_done = bar0 + 8 * TILES + 8 * slot;
        uint8_t* a_tile = sm + L::A + slot * (TM * 128);
        const float* bm = reinterpret_cast<const float*>(sm + L::BM);
        const float* bp = reinterpret_cast<const float*>(sm + L::BP);
        const float* wo = reinterpret_cast<const float*>(sm + L::WO);
        const float* woc = reinterpret_cast<const float*>(sm + L::WOC);
        const float* boc = reinterpret_cast<const float*>(sm + L::BOC);
        float* xch = reinterpret_cast<float*>(sm + L::XCH) + slot * TM * 4;
        uint32_t step = 0;   // handshakes completed by this slot: parity of both barriers
        for (int tile = cta * TILES + slot; tile < ntiles; tile += ncta * TILES) {
            const int s = tile * TM + row;
            bool active = s < P.P;
            float p[3] = {0.f, 0.f, 0.f};
            if (active) {
                if (P.pts) { p[0] = P.pts[3 * (size_t)s]; p[1] = P.pts[3 * (size_t)s + 1]; p[2] = P.pts[3 * (size_t)s + 2]; }
                else {
                    const int ray = s / P.S;
                    const uint8_t ok = P.valid ? P.valid[ray] : (uint8_t)1;
                    const float z = P.z[s];
#pragma unroll
                    for (int a = 0; a < 3; ++a) p[a] = __fadd_rn(P.rays_o[3 * ray + a], __fmul_rn(P.rays_d[3 * ray + a], z));   // Renderer.cpp:121
                    if (!ok) active = false;
                }
            }
            // ---- grid features, quad-cooperative: in pass r the quad q of this warp serves sample 8r + q of the warp
            float cmid[C == 64 ? 4 : 1][8];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int j = 8 * r + q;
                float pj[3];
#pragma unroll
                for (int a = 0; a < 3; ++a) pj[a] = __shfl_sync(0xffffffffu, p[a], j);
                const bool act = __shfl_sync(0xffffffffu, active ? 1 : 0, j) != 0;
                float c8[8], oc[NO];
#pragma unroll
                for (int o = 0; o < NO; ++o) oc[o] = 0.0f;
                if (act) gather8(P.grid[dec], P.bnd, pj, t, c8);
                else {
#pragma unroll
                    for (int i = 0; i < 8; ++i) c8[i] = 0.0f;
                }
#pragma unroll
                for (int o = 0; o < NO; ++o)
#pragma unroll
                    for (int i = 0; i < 8; ++i) oc[o] = fmaf(c8[i], woc[o * C + 8 * t + i], oc[o]);
                if (C == 64) {   // fine decoder: cat(fine, middle) (MLP.cpp:79-84); the middle half goes out as the second chunk
                    float* cm = cmid[C == 64 ? r : 0];
                    if (act) gather8(P.grid[1], P.bnd, pj, t, cm);
                    else {
#pragma unroll
                        for (int i = 0; i < 8; ++i) cm[i] = 0.0f;
                    }
#pragma unroll
                    for (int o = 0; o < NO; ++o)
#pragma unroll
                        for (int i = 0; i < 8; ++i) oc[o] = fmaf(cm[i], woc[o * C + 32 + 8 * t + i], oc[o]);
                }
#pragma unroll
                for (int o = 0; o < NO; ++o) oc[o] = quad_sum(oc[o]);
                const int rj = (wq << 5) | j;
                if (t == 0) {
#pragma unroll
                    for (int o = 0; o < NO; ++o) xch[rj * 4 + o] = oc[o] + boc[o];
                }
                uint4 hi, lo;
                split8(c8, hi, lo);
                *reinterpret_cast<uint4*>(a_tile + chunk_off(rj, t)) = hi;
                *reinterpret_cast<uint4*>(a_tile + chunk_off(rj, 4 + t)) = lo;
            }
            __syncwarp();         // xch rows written by the quads are read by their owner lanes at the output layer
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            fence_before();
            mbar_arrive(full_a);
            ++step;
            if (C == 64) {
                mbar_wait(mma_done, (step - 1) & 1);
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const int rj = (wq << 5) | (8 * r + q);
                    uint4 hi, lo;
                    split8(cmid[C == 64 ? r : 0], hi, lo);
                    *reinterpret_cast<uint4*>(a_tile + chunk_off(rj, t)) = hi;
                    *reinterpret_cast<uint4*>(a_tile + chunk_off(rj, 4 + t)) = lo;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                fence_before();
                mbar_arrive(full_a);
                ++step;
            }
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
                if (!active) {
#pragma unroll
                    for (int k = 0; k < 32; ++k) e[k] = 0.0f;
                }
                mbar_wait(mma_done, (step - 1) & 1);       // the previous chunk's MMAs have consumed the A tile
                store_row32(a_tile, row, e);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                fence_before();
                mbar_arrive(full_a);
                ++step;
            }
            // ---- five layers: read the accumulator, bias + relu (+ mask), hand u_i back as the next A tile
#pragma unroll 1
            for (int i = 0; i < 5; ++i) {
                mbar_wait(mma_done, (step - 1) & 1);
                fence_after();
                float v[32];
                tmem_ld32(tm + (i == 0 ? 0 : i == 3 ? 32 : i == 1 ? 64 : i == 2 ? 96 : 128), v);
                tmem_ld_wait();
                uint32_t mask = 0;
#pragma unroll
                for (int k4 = 0; k4 < 8; ++k4) {
                    const float4 b4 = *reinterpret_cast<const float4*>(bp + i * HID + 4 * k4);
                    const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        const float a = v[4 * k4 + r] + bb[r];
                        const bool pos = a > 0.0f;
                        mask |= (pos ? 1u : 0u) << (4 * k4 + r);
                        v[4 * k4 + r] = pos ? a : 0.0f;
                    }
                }
                if (P.masks && active) P.masks[((size_t)(dec - 1) * 5 + i) * P.mask_stride + s] = mask;
                if (i < 4) {
                    store_row32(a_tile, row, v);
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    fence_before();
                    mbar_arrive(full_a);
                    ++step;
                } else {
                    float out[NO];
#pragma unroll
                    for (int o = 0; o < NO; ++o) {
                        float acc = xch[row * 4 + o];          // written by this warp's quads before the first arrive of the tile
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
            __syncwarp();         // xch rows of this warp are rewritten by the next tile's gather
            fence_before();       // the accumulator reads above are ordered before the next tile's first arrive
        }
    } else if (lane == 0) {
        // ------------------------------------------------------------------ MMA issuer of one tile slot
        const int slot = warp - TILES * 4;
        const uint32_t tm = tmem + slot * SLOT_COLS;
        const uint32_t full_a = bar0 + 8 * slot, mma_done = bar0 + 8 * TILES + 8 * slot;
        const uint32_t sbase = smem_u32(sm);
        const uint64_t a = make_desc(sbase + L::A + slot * (TM * 128));
        constexpr uint32_t I32 = make_idesc(128, 32), I64 = make_idesc(128, 64), I160 = make_idesc(128, 160);
        uint32_t step = 0;
        for (int tile = cta * TILES + slot; tile < ntiles; tile += ncta * TILES) {
            for (int cc = 0; cc < C / 32; ++cc) {   // all grid-feature terms at once: [acc0 = 0 | acc3 | acc1 | acc2 | acc4] (+)= c [0; G_2; G_0; G_1; G_3]^T
                mbar_wait(full_a, step & 1); fence_after();
                issue3(tm, a, make_desc(sbase + L::GC + cc * 20480), I160, cc == 0);
                mma_commit(mma_done); ++step;
            }
            for (int j = 0; j < 3; ++j) {           // [acc0 | acc3] += e_j [W0 ; W3E]_j^T
                mbar_wait(full_a, step & 1); fence_after();
                issue3(tm, a, make_desc(sbase + L::WE + j * 8192), I64, false);
                mma_commit(mma_done); ++step;
            }
            for (int l = 0; l < 4; ++l) {           // acc_{l+1} += u_l W_{l+1}^T
                mbar_wait(full_a, step & 1); fence_after();
                issue3(tm + acc_col(l + 1), a, make_desc(sbase + L::WH + l * 4096), I32, false);
                mma_commit(mma_done); ++step;
            }
        }
    }
    fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u));
}

__global__ void __launch_bounds__(THREADS, 1) k_decode_fwd_tc16(const DecodeParams P) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    int dec = 1;
#pragma unroll
    for (int d = 2; d < 4; ++d) if ((int)blockIdx.x >= P.cta_begin[d]) dec = d;
    const int cta = blockIdx.x - P.cta_begin[dec], ncta = P.cta_begin[dec + 1] - P.cta_begin[dec];
    if (dec == 1) run_decoder<32, 1>(P, sm, 1, cta, ncta);
    else if (dec == 2) run_decoder<64, 1>(P, sm, 2, cta, ncta);
    else run_decoder<32, 4>(P, sm, 3, cta, ncta);
}

}  // namespace tc16

cudaError_t launch_decode_fwd_tc16(const DecodeParams& P, int grid, cudaStream_t st) {
    const size_t smem = (size_t)tc16::Smem<64>::TOTAL + 1024;
    static unsigned attr_done = 0;      // per device: function attributes are per-device state
    int dev = 0; cudaGetDevice(&dev);
    if (!((attr_done >> (dev & 31)) & 1u)) {
        cudaError_t e = cudaFuncSetAttribute(tc16::k_decode_fwd_tc16, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attr_done |= 1u << (dev & 31);
    }
    tc16::k_decode_fwd_tc16<<<grid, tc16::THREADS, smem, st>>>(P);
    return cudaGetLastError();
}

}  // namespace nsb
