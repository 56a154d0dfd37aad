// Forward decoder kernel on the 5th-generation tensor cores (tcgen05.mma kind::tf32, accumulators in tensor memory).
//
// One CTA per SM keeps ONE decoder's weights resident in shared memory in the UMMA canonical K-major SWIZZLE_128B layout
// (tf32 hi plane + residual lo plane for the fp32-grade 3xTF32 product) and runs two 128-sample tiles concurrently:
//   warps 0-7 / 8-15  TWO threads per sample (each owns 16 of the 32 columns): trilinear gather, Fourier features,
//                     per-layer epilogues (bias, relu, mask, hi/lo split) -- activations go back to the tensor core through a 128x32 shared-memory A tile
//                     (written with 16-byte swizzled stores), the grid features through tensor memory (A-from-TMEM MMA)
//   warps 16 / 17     one elected lane each issues the tcgen05.mma stream of its tile and commits to an mbarrier
// The two tiles ping-pong: while one tile's threads run an epilogue the other tile's MMAs execute.
//
// Algebra used to keep the chain to ONE accumulator read per layer: with h_{i+1} = relu(a_i) + Fc_i c + bc_i and
// a_{i+1} = W_{i+1} h_{i+1} + b_{i+1}, the grid-feature term is folded into the NEXT layer's accumulator with composed
// weights  G_i = W_{i+1} Fc_i,  b'_{i+1} = b_{i+1} + W_{i+1} bc_i  (built by k_compose), so that
//   a_{i+1} = W_{i+1} relu(a_i) + G_i c + b'_{i+1}
// and the A operand of every layer is just u_i = relu(a_i).  The output layer becomes Wo u_4 + (Wo Fc_4) c + const.
// Replaces NICE::forward / MLP::forward (NICE.cpp:16-51, MLP.cpp:76-102) like decode_fwd.cu; selected by the host when no
// weight-gradient stash is requested.
#include "decode.cuh"
#include "params.h"

namespace nsb {
namespace tc {

#ifdef NSB_TC_TIMING
#define TC_T0() long long t_ = clock64()
#define TC_ACC(k) do { const long long n_ = clock64(); tacc[k] += n_ - t_; t_ = n_; } while (0)
#else
#define TC_T0()
#define TC_ACC(k)
#endif

constexpr int TM = 128;                       // samples per tile (UMMA M)
constexpr int GROUPS = 2;                     // tiles in flight per CTA
constexpr int GTHREADS = 256;                 // compute threads per tile: TWO per sample, each owns 16 of the 32 columns
constexpr int CTHREADS = GROUPS * GTHREADS;   // compute threads (warps 0..15)
constexpr int THREADS = CTHREADS + 32 * GROUPS;   // + one issuer warp per tile group

// ---- composed weights (global, per decoder): G[4][32][C] | bp[5][32] | woc[4][C] | boc[4] ---------------------------
__host__ __device__ constexpr int comp_G(int) { return 0; }
__host__ __device__ constexpr int comp_bp(int C) { return 4 * HID * C; }
__host__ __device__ constexpr int comp_woc(int C) { return comp_bp(C) + 5 * HID; }
__host__ __device__ constexpr int comp_boc(int C) { return comp_woc(C) + 4 * C; }
__host__ __device__ constexpr int comp_total(int C) { return comp_boc(C) + 4; }

// one block per (matrix i = 0..3 | 4 = output), 256 threads
template <int C, int O>
__device__ void compose_one(const float* __restrict__ flat, float* __restrict__ comp, int i, int tid, int nthr) {
    const DecFlat f = DecFlat::make(C, O);
    if (i < 4) {
        // W_{i+1} (for i == 2 the hidden columns of the skip layer), row stride ldw
        const float* Wn = flat + f.W[i + 1] + (i == 2 ? EMB : 0);
        const int ldw = i == 2 ? EMB + HID : HID;
        const float* Fc = flat + f.Fc[i];
        const float* bc = flat + f.bc[i];
        for (int idx = tid; idx < HID * C; idx += nthr) {
            const int o = idx / C, ch = idx % C;
            float s = 0.0f;
            for (int m = 0; m < HID; ++m) s = fmaf(Wn[o * ldw + m], Fc[m * C + ch], s);
            comp[comp_G(C) + i * HID * C + idx] = s;
        }
        for (int o = tid; o < HID; o += nthr) {
            float s = flat[f.b[i + 1] + o];
            for (int m = 0; m < HID; ++m) s = fmaf(Wn[o * ldw + m], bc[m], s);
            comp[comp_bp(C) + (i + 1) * HID + o] = s;
            if (i == 0) comp[comp_bp(C) + o] = flat[f.b[0] + o];
        }
    } else {
        const float* Wo = flat + f.Wo;
        const float* Fc = flat + f.Fc[4];
        const float* bc = flat + f.bc[4];
        for (int idx = tid; idx < 4 * C; idx += nthr) {
            const int o = idx / C, ch = idx % C;
            float s = 0.0f;
            if (o < O) for (int m = 0; m < HID; ++m) s = fmaf(Wo[o * HID + m], Fc[m * C + ch], s);
            comp[comp_woc(C) + idx] = s;
        }
        if (tid < 4) {
            float s = 0.0f;
            if (tid < O) { s = flat[f.bo + tid]; for (int m = 0; m < HID; ++m) s = fmaf(Wo[tid * HID + m], bc[m], s); }
            comp[comp_boc(C) + tid] = s;
        }
    }
}

struct ComposeParams { const float* flat[4]; float* comp[4]; int mask; };   // mask: bit d set -> compose decoder d (1..3)

__global__ void k_compose(ComposeParams P) {
    const int dec = 1 + blockIdx.x / 5, i = blockIdx.x % 5;
    if (!((P.mask >> dec) & 1)) return;
    if (dec == 1) compose_one<32, 1>(P.flat[1], P.comp[1], i, threadIdx.x, blockDim.x);
    else if (dec == 2) compose_one<64, 1>(P.flat[2], P.comp[2], i, threadIdx.x, blockDim.x);
    else compose_one<32, 4>(P.flat[3], P.comp[3], i, threadIdx.x, blockDim.x);
}

// ---- shared-memory image (bytes).  Every UMMA tile starts on a multiple of 1024 B. --------------------------------------
template <int C>
struct Smem {
    static constexpr int WE = 0;                                  // 3 chunks x { [64][32] hi 8 KB | lo 8 KB }: rows 0-31 W0, 32-63 W3E
    static constexpr int WH = WE + 3 * 16384;                     // 4 layers x { [32][32] hi 4 KB | lo 4 KB }: W1, W2, W3H, W4
    static constexpr int G = WH + 4 * 8192;                       // 4 x (C/32) x { hi 4 KB | lo 4 KB }
    static constexpr int A = G + 4 * (C / 32) * 8192;             // GROUPS x { [128][32] hi 16 KB | lo 16 KB }
    static constexpr int BM = A + GROUPS * 32768;                 // Fourier matrix [3][96] floats
    static constexpr int BP = BM + 3 * EMBP * 4;                  // b'[5][32]
    static constexpr int WO = BP + 5 * HID * 4;                   // Wo[4][32]
    static constexpr int WOC = WO + 4 * HID * 4;                  // (Wo Fc_4)[4][C]
    static constexpr int BOC = WOC + 4 * C * 4;                   // const[4]
    static constexpr int BAR = BOC + 16;                          // mbarriers: full_A[2], mma_done[2]
    static constexpr int TMEMPTR = BAR + 4 * 8;
    static constexpr int XCH = TMEMPTR + 16;                      // GROUPS x 128 x float4: output partial sums of column half 1
    static constexpr int TOTAL = XCH + GROUPS * 128 * 16;
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
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
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
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]),
                   "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float* v) {
    const uint32_t* r = reinterpret_cast<const uint32_t*>(v);
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]),
                   "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float* v) {
    const uint32_t* r = reinterpret_cast<const uint32_t*>(v);
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]),
                   "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]),
                   "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// float offset of element (row r, k) of a [rows][32] fp32 tile in the K-major SWIZZLE_128B layout (8-row atoms of 1 KB)
__device__ __forceinline__ int sw128(int r, int k) { return (r >> 3) * 256 + (r & 7) * 32 + (((k >> 2) ^ (r & 7)) << 2) + (k & 3); }

__device__ __forceinline__ void put_hl(float* hi_tile, int lo_off_floats, int idx, float v) {
    const float h = __uint_as_float(f2tf32(v));
    hi_tile[idx] = h;
    hi_tile[idx + lo_off_floats] = v - h;
}

template <int C, int O>
__device__ void stage(uint8_t* sm, const float* __restrict__ flat, const float* __restrict__ comp, int tid, int nthr) {
    using L = Smem<C>;
    const DecFlat f = DecFlat::make(C, O);
    float* we = reinterpret_cast<float*>(sm + L::WE);
    for (int idx = tid; idx < 3 * 64 * 32; idx += nthr) {
        const int j = idx / 2048, r = (idx / 32) % 64, k = idx % 32, ft = 32 * j + k;
        float v = 0.0f;
        if (ft < EMB) v = r < 32 ? flat[f.W[0] + r * EMB + ft] : flat[f.W[3] + (r - 32) * (EMB + HID) + ft];
        put_hl(we + j * 4096, 2048, sw128(r, k), v);
    }
    float* wh = reinterpret_cast<float*>(sm + L::WH);
    for (int idx = tid; idx < 4 * 32 * 32; idx += nthr) {
        const int l = idx / 1024, o = (idx / 32) % 32, k = idx % 32;   // l = 0..3 <-> layers 1..4
        const float v = l == 2 ? flat[f.W[3] + o * (EMB + HID) + EMB + k] : flat[f.W[l + 1] + o * HID + k];
        put_hl(wh + l * 2048, 1024, sw128(o, k), v);
    }
    float* gg = reinterpret_cast<float*>(sm + L::G);
    for (int idx = tid; idx < 4 * HID * C; idx += nthr) {
        const int i = idx / (HID * C), o = (idx / C) % HID, ch = idx % C;
        put_hl(gg + (i * (C / 32) + ch / 32) * 2048, 1024, sw128(o, ch % 32), comp[comp_G(C) + idx]);
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

// 3xTF32 product over one K = 32 block (4 UMMA k-steps of 8): D += A_lo B_hi + A_hi B_lo + A_hi B_hi
__device__ __forceinline__ void issue_ss(uint32_t d, uint64_t a_hi, uint64_t a_lo, uint64_t b_hi, uint64_t b_lo, uint32_t idesc, bool zero_first) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        mma_ss(d, a_lo + 2 * k, b_hi + 2 * k, idesc, (zero_first && k == 0) ? 0u : 1u);
        mma_ss(d, a_hi + 2 * k, b_lo + 2 * k, idesc, 1u);
        mma_ss(d, a_hi + 2 * k, b_hi + 2 * k, idesc, 1u);
    }
}
__device__ __forceinline__ void issue_ts(uint32_t d, uint32_t a_hi, uint32_t a_lo, uint64_t b_hi, uint64_t b_lo, uint32_t idesc, bool zero_first) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        mma_ts(d, a_lo + 8 * k, b_hi + 2 * k, idesc, (zero_first && k == 0) ? 0u : 1u);
        mma_ts(d, a_hi + 8 * k, b_lo + 2 * k, idesc, 1u);
        mma_ts(d, a_hi + 8 * k, b_hi + 2 * k, idesc, 1u);
    }
}

// tensor-memory columns of one tile group
constexpr int ACC0 = 0, ACCS = 32, ACC1 = 64, ACH = 96;
template <int C> constexpr int acl() { return ACH + C; }
constexpr int GROUP_COLS = 256;

// the thread's sample -> 16 channels [16h, 16h+16) of the trilinear feature (8 corners x 4 float4)
__device__ __forceinline__ void gather16(const GridView& G, const Bound& bnd, const float (&p)[3], int h, float (&c)[16]) {
    Tri s;
    tri_setup(G, bnd, p, s);
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = 0.0f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        int off;
        const float w = tri_corner(G, s, k, off);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float4 v = ldg4(G.data + off + 16 * h + 4 * q);
            c[4 * q] = fmaf(v.x, w, c[4 * q]); c[4 * q + 1] = fmaf(v.y, w, c[4 * q + 1]);
            c[4 * q + 2] = fmaf(v.z, w, c[4 * q + 2]); c[4 * q + 3] = fmaf(v.w, w, c[4 * q + 3]);
        }
    }
}

// write columns [16h, 16h+16) of the thread's row of the A tile: hi plane at `hi_tile`, lo plane 4096 floats further
__device__ __forceinline__ void store_half(float* hi_tile, int row, int h, const float (&v)[16]) {
    float* base = hi_tile + (row >> 3) * 256 + (row & 7) * 32;
    const int sw = row & 7;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        float4 hh, l;
        hh.x = __uint_as_float(f2tf32(v[4 * c])); hh.y = __uint_as_float(f2tf32(v[4 * c + 1]));
        hh.z = __uint_as_float(f2tf32(v[4 * c + 2])); hh.w = __uint_as_float(f2tf32(v[4 * c + 3]));
        l.x = v[4 * c] - hh.x; l.y = v[4 * c + 1] - hh.y; l.z = v[4 * c + 2] - hh.z; l.w = v[4 * c + 3] - hh.w;
        float* d = base + (((4 * h + c) ^ sw) << 2);
        *reinterpret_cast<float4*>(d) = hh;
        *reinterpret_cast<float4*>(d + 4096) = l;
    }
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
        for (int g = 0; g < GROUPS; ++g) { mbar_init(bar0 + 8 * g, GTHREADS); mbar_init(bar0 + 16 + 8 * g, 1); }
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (dec == 1) stage<32, 1>(sm, P.dec_flat[1], P.comp[1], tid, THREADS);
    else if (dec == 2) stage<64, 1>(sm, P.dec_flat[2], P.comp[2], tid, THREADS);
    else stage<32, 4>(sm, P.dec_flat[3], P.comp[3], tid, THREADS);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = *tmem_ptr;

    if (warp < GROUPS * 8) {
        // ---------------------------------------------- compute threads: two per sample (h = column half), warp w -> TMEM lanes 32 (w & 3)
        const int grp = warp >> 3, h = (warp >> 2) & 1, tg = ((warp & 3) << 5) | lane;
        const uint32_t tm = tmem + grp * GROUP_COLS + ((uint32_t)((warp & 3) * 32) << 16);
        const uint32_t full_a = bar0 + 8 * grp, mma_done = bar0 + 16 + 8 * grp;
        float* a_hi = reinterpret_cast<float*>(sm + L::A + grp * 32768);
        const float* bm = reinterpret_cast<const float*>(sm + L::BM);
        const float* bp = reinterpret_cast<const float*>(sm + L::BP);
        const float* wo = reinterpret_cast<const float*>(sm + L::WO);
        const float* woc = reinterpret_cast<const float*>(sm + L::WOC);
        const float* boc = reinterpret_cast<const float*>(sm + L::BOC);
        float* xch = reinterpret_cast<float*>(sm + L::XCH) + grp * 128 * 4;    // half-1 -> half-0 exchange of the output partial sums
        constexpr int NO = O == 4 ? 3 : 1;
        uint32_t step = 0;   // handshakes completed by this group: parity of both barriers
#ifdef NSB_TC_TIMING
        long long tacc[6] = {0, 0, 0, 0, 0, 0};   // gather, sin, wait(E), wait(layers), epilogue, tiles
#endif
        for (int tile = cta * GROUPS + grp; tile < ntiles; tile += ncta * GROUPS) {
            TC_T0();
            const int s = tile * TM + tg;
            bool active = s < P.P;
            float p[3] = {0.f, 0.f, 0.f};
            if (active) {
                if (P.pts) { p[0] = P.pts[3 * (size_t)s]; p[1] = P.pts[3 * (size_t)s + 1]; p[2] = P.pts[3 * (size_t)s + 2]; }
                else {
                    const int ray = s / P.S;
                    if (P.valid && !P.valid[ray]) active = false;
                    else {
                        const float z = P.z[s];
#pragma unroll
                        for (int a = 0; a < 3; ++a) p[a] = __fadd_rn(P.rays_o[3 * ray + a], __fmul_rn(P.rays_d[3 * ray + a], z));   // Renderer.cpp:121
                    }
                }
            }
            // grid features -> tensor memory (A operand of the G_i c products) and this half's share of the output-layer constant
            float outc[NO];
#pragma unroll
            for (int o = 0; o < NO; ++o) outc[o] = h == 0 ? boc[o] : 0.0f;
#pragma unroll
            for (int cc = 0; cc < C / 32; ++cc) {
                float c[16];
                if (active) gather16(P.grid[cc == 0 ? dec : 1], P.bnd, p, h, c);   // fine decoder: cat(fine, middle) (MLP.cpp:79-84)
                else {
#pragma unroll
                    for (int i = 0; i < 16; ++i) c[i] = 0.0f;
                }
#pragma unroll
                for (int o = 0; o < NO; ++o) {
                    float acc = outc[o];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float4 w4 = *reinterpret_cast<const float4*>(woc + o * C + 32 * cc + 16 * h + 4 * q);
                        acc = fmaf(c[4 * q], w4.x, acc); acc = fmaf(c[4 * q + 1], w4.y, acc); acc = fmaf(c[4 * q + 2], w4.z, acc); acc = fmaf(c[4 * q + 3], w4.w, acc);
                    }
                    outc[o] = acc;
                }
                float lo[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) { const float hh = __uint_as_float(f2tf32(c[i])); lo[i] = c[i] - hh; c[i] = hh; }
                tmem_st16(tm + ACH + 32 * cc + 16 * h, c);
                tmem_st16(tm + acl<C>() + 32 * cc + 16 * h, lo);
            }
            tmem_st_wait();
            TC_ACC(0);
            // Fourier features, 32 per handshake (16 per thread), through the shared-memory A tile
#pragma unroll 1
            for (int j = 0; j < 3; ++j) {
                float e[16];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int ft = 32 * j + 16 * h + 4 * q;
                    const float4 b0 = *reinterpret_cast<const float4*>(bm + ft), b1 = *reinterpret_cast<const float4*>(bm + EMBP + ft), b2 = *reinterpret_cast<const float4*>(bm + 2 * EMBP + ft);
                    e[4 * q] = ff_sin(fmaf(p[2], b2.x, fmaf(p[1], b1.x, p[0] * b0.x)));
                    e[4 * q + 1] = ff_sin(fmaf(p[2], b2.y, fmaf(p[1], b1.y, p[0] * b0.y)));
                    e[4 * q + 2] = ff_sin(fmaf(p[2], b2.z, fmaf(p[1], b1.z, p[0] * b0.z)));
                    e[4 * q + 3] = ff_sin(fmaf(p[2], b2.w, fmaf(p[1], b1.w, p[0] * b0.w)));   // padded columns of B are 0 -> sin(0) = 0
                }
                if (!active) {
#pragma unroll
                    for (int k = 0; k < 16; ++k) e[k] = 0.0f;
                }
                TC_ACC(1);
                if (j > 0) { mbar_wait(mma_done, (step - 1) & 1); }     // previous chunk's MMAs have consumed the A tile
                TC_ACC(2);
                store_half(a_hi, tg, h, e);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                fence_before();
                mbar_arrive(full_a);
                ++step;
            }
            // five layers: read the accumulator, bias + relu (+ mask), hand u_i back as the next A tile
#pragma unroll 1
            for (int i = 0; i < 5; ++i) {
                TC_ACC(4);
                mbar_wait(mma_done, (step - 1) & 1);
                TC_ACC(3);
                fence_after();
                float v[16];
                tmem_ld16(tm + ((i & 1) ? ACC1 : ACC0) + 16 * h, v);
                if (i == 3) {
                    float sk[16];
                    tmem_ld16(tm + ACCS + 16 * h, sk);
                    tmem_ld_wait();
#pragma unroll
                    for (int k = 0; k < 16; ++k) v[k] += sk[k];
                } else {
                    tmem_ld_wait();
                }
                uint32_t mask = 0;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 b4 = *reinterpret_cast<const float4*>(bp + i * HID + 16 * h + 4 * q);
                    const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        const float a = v[4 * q + r] + bb[r];
                        const bool pos = a > 0.0f;
                        mask |= (pos ? 1u : 0u) << (4 * q + r);
                        v[4 * q + r] = pos ? a : 0.0f;
                    }
                }
                if (P.masks && active)
                    reinterpret_cast<uint16_t*>(P.masks + ((size_t)(dec - 1) * 5 + i) * P.mask_stride + s)[h] = (uint16_t)mask;
                if (i < 4) {
                    store_half(a_hi, tg, h, v);
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    fence_before();
                    mbar_arrive(full_a);
                    ++step;
                } else {
                    float out[NO];
#pragma unroll
                    for (int o = 0; o < NO; ++o) {
                        float acc = outc[o];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const float4 w4 = *reinterpret_cast<const float4*>(wo + o * HID + 16 * h + 4 * q);
                            acc = fmaf(v[4 * q], w4.x, acc); acc = fmaf(v[4 * q + 1], w4.y, acc); acc = fmaf(v[4 * q + 2], w4.z, acc); acc = fmaf(v[4 * q + 3], w4.w, acc);
                        }
                        out[o] = acc;
                    }
                    // the two halves of a row live in different warps: half 1 passes its partial sums through shared memory
                    if (h == 1) {
#pragma unroll
                        for (int o = 0; o < NO; ++o) xch[tg * 4 + o] = out[o];
                    }
                    asm volatile("bar.sync %0, %1;" ::"r"(1 + grp), "r"(GTHREADS) : "memory");
                    if (h == 0 && active) {
#pragma unroll
                        for (int o = 0; o < NO; ++o) out[o] += xch[tg * 4 + o];
                        if (O == 4) *reinterpret_cast<float4*>(P.out_rgb + 4 * (size_t)s) = make_float4(out[0], out[NO > 1 ? 1 : 0], out[NO > 2 ? 2 : 0], 0.0f);
                        else P.out_occ[dec][s] = out[0];
                    }
                    asm volatile("bar.sync %0, %1;" ::"r"(1 + grp), "r"(GTHREADS) : "memory");   // xch is reused by the next tile
                }
            }
            fence_before();   // the accumulator reads above are ordered before the next tile's first arrive
            TC_ACC(4);
#ifdef NSB_TC_TIMING
            tacc[5] += 1;
#endif
        }
#ifdef NSB_TC_TIMING
        if (P.dbg && tg == 0 && h == 0) for (int k = 0; k < 6; ++k) atomicAdd(P.dbg + 8 * dec + k, (unsigned long long)tacc[k]);
#endif
    } else if (lane == 0) {
        // ------------------------------------------------------------------ MMA issuer of tile group (warp - 8)
        const int grp = warp - GROUPS * 8;
        const uint32_t tm = tmem + grp * GROUP_COLS;
        const uint32_t full_a = bar0 + 8 * grp, mma_done = bar0 + 16 + 8 * grp;
        const uint32_t sbase = smem_u32(sm);
        const uint64_t a_hi = make_desc(sbase + L::A + grp * 32768), a_lo = make_desc(sbase + L::A + grp * 32768 + 16384);
        constexpr uint32_t I32 = make_idesc(128, 32), I64 = make_idesc(128, 64);
        uint32_t step = 0;
        auto cterm = [&](int i, int dcol) {   // D = G_i c  (A from tensor memory)
#pragma unroll
            for (int cc = 0; cc < C / 32; ++cc) {
                const uint32_t g = sbase + L::G + (i * (C / 32) + cc) * 8192;
                issue_ts(tm + dcol, tm + ACH + 32 * cc, tm + acl<C>() + 32 * cc, make_desc(g), make_desc(g + 4096), I32, cc == 0);
            }
        };
        for (int tile = cta * GROUPS + grp; tile < ntiles; tile += ncta * GROUPS) {
            for (int j = 0; j < 3; ++j) {      // [acc0 | accS] = e [W0 ; W3E]^T, one 32-feature chunk per handshake
                mbar_wait(full_a, step & 1); fence_after();
                const uint32_t w = sbase + L::WE + j * 16384;
                issue_ss(tm + ACC0, a_hi, a_lo, make_desc(w), make_desc(w + 8192), I64, j == 0);
                mma_commit(mma_done); ++step;
            }
            cterm(0, ACC1);
            for (int l = 0; l < 4; ++l) {      // layers 1..4: acc[(l+1)&1] += W_{l+1} u_l ; then pre-load the next c-term
                mbar_wait(full_a, step & 1); fence_after();
                const uint32_t w = sbase + L::WH + l * 8192;
                issue_ss(tm + ((l & 1) ? ACC0 : ACC1), a_hi, a_lo, make_desc(w), make_desc(w + 4096), I32, false);
                mma_commit(mma_done); ++step;
                if (l < 3) cterm(l + 1, (l & 1) ? ACC1 : ACC0);
            }
        }
    }
    fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u));
}

__global__ void __launch_bounds__(THREADS, 1) k_decode_fwd_tc(const DecodeParams P) {
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

}  // namespace tc

size_t decode_fwd_tc_smem() { return (size_t)tc::Smem<64>::TOTAL + 1024; }

cudaError_t launch_compose(const float* const flat[4], float* const comp[4], int mask, cudaStream_t st) {
    tc::ComposeParams C;
    for (int d = 0; d < 4; ++d) { C.flat[d] = flat[d]; C.comp[d] = comp[d]; }
    C.mask = mask;
    tc::k_compose<<<15, 256, 0, st>>>(C);
    return cudaGetLastError();
}
int compose_floats(int which) { return which == 2 ? tc::comp_total(64) : tc::comp_total(32); }

cudaError_t launch_decode_fwd_tc(const DecodeParams& P, int grid, cudaStream_t st) {
    const size_t smem = decode_fwd_tc_smem();
    static unsigned attr_done = 0;      // per device: function attributes are per-device state
    int dev = 0; cudaGetDevice(&dev);
    if (!((attr_done >> (dev & 31)) & 1u)) {
        cudaError_t e = cudaFuncSetAttribute(tc::k_decode_fwd_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attr_done |= 1u << (dev & 31);
    }
    tc::k_decode_fwd_tc<<<grid, tc::THREADS, smem, st>>>(P);
    return cudaGetLastError();
}

}  // namespace nsb
