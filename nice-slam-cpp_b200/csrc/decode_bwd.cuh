// Backward decoder kernel: reads the relu bit masks the training forward saved for each 16-sample tile (no
// forward recomputation), back-propagates the raw cotangent through the MLP on tensor cores, scatters the grid-feature
// gradient with 128-bit vector reductions into the channel-last gradient grids, accumulates the ray (pose)
// gradient, and -- for the colour decoder when its weights are being optimised -- stashes the per-layer
// activations / gradients for the split-K weight-gradient kernel (wgrad.cu).
// This is the autograd backward of NICE::forward (NICE.cpp:43-50) that loss.backward() runs at
// Mapper.cpp:444 / Tracker.cpp:84.
#pragma once
#include "decode.cuh"
#include "params.h"

namespace nsb {

enum { F_GRID = 1, F_WGRAD = 2, F_RAY = 4 };

#ifndef NSB_SCATTER_AGG
#define NSB_SCATTER_AGG 1   // warp-aggregated grid-gradient scatter (scatter_tile); 0 = per-sample quad reductions (grid_backward)
#endif
#ifndef NSB_BWD_MIN_CTAS
#define NSB_BWD_MIN_CTAS 1   // 512 threads x 128 registers: one CTA of 16 warps per SM
#endif

__device__ __forceinline__ void zero_tile(float (&a)[4][4]) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int q = 0; q < 4; ++q) a[j][q] = 0.0f;
}
__device__ __forceinline__ void apply_mask(float (&gu)[4][4], const float (&gh)[4][4], uint32_t m) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int q = 0; q < 4; ++q) gu[j][q] = ((m >> (4 * j + q)) & 1u) ? gh[j][q] : 0.0f;
}

// Scatter the feature gradient of the thread's two samples into grid G (channels 4t..4t+3 and 16+4t..16+4t+3), and/or
// accumulate d c / d p (coordinate derivative of the trilinear sample) into gp.  gc[r][8].
template <bool DO_GRID, bool DO_RAY>
__device__ __forceinline__ void grid_backward(const GridView& G, const Bound& bnd, const float (&p)[2][3],
                                              const float (&gc)[2][8], int t, float (&gp)[2][3]) {
    Tri s[2];
    tri_setup(G, bnd, p[0], s[0]);
    tri_setup(G, bnd, p[1], s[1]);
    if (DO_RAY) {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            float gx = 0.0f, gy = 0.0f, gz = 0.0f;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                int off;
                tri_corner(G, s[r], k, off);
                const float4 v0 = ldg4(G.data + off + 4 * t), v1 = ldg4(G.data + off + 16 + 4 * t);
                const float dot = gc[r][0] * v0.x + gc[r][1] * v0.y + gc[r][2] * v0.z + gc[r][3] * v0.w +
                                  gc[r][4] * v1.x + gc[r][5] * v1.y + gc[r][6] * v1.z + gc[r][7] * v1.w;
                const int dx = k & 1, dy = (k >> 1) & 1, dz = (k >> 2) & 1;
                const float wx = dx ? s[r].w1[0] : s[r].w0[0], wy = dy ? s[r].w1[1] : s[r].w0[1], wz = dz ? s[r].w1[2] : s[r].w0[2];
                gx += (dx ? dot : -dot) * wy * wz;
                gy += (dy ? dot : -dot) * wx * wz;
                gz += (dz ? dot : -dot) * wx * wy;
            }
            gp[r][0] += gx * s[r].gm[0]; gp[r][1] += gy * s[r].gm[1]; gp[r][2] += gz * s[r].gm[2];   // partial over the quad's channels
        }
    }
    if (DO_GRID) {
        // the thread's two samples are neighbours on the ray: when they fall in the same cell one reduction serves both
        const bool same = s[0].i0[0] == s[1].i0[0] && s[0].i0[1] == s[1].i0[1] && s[0].i0[2] == s[1].i0[2];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            int off0, off1;
            const float w0 = tri_corner(G, s[0], k, off0);
            const float w1 = tri_corner(G, s[1], k, off1);
            float* a0 = G.grad + off0 + 4 * t;
            if (same) {
                red_add_v4(a0, w0 * gc[0][0] + w1 * gc[1][0], w0 * gc[0][1] + w1 * gc[1][1], w0 * gc[0][2] + w1 * gc[1][2], w0 * gc[0][3] + w1 * gc[1][3]);
                red_add_v4(a0 + 16, w0 * gc[0][4] + w1 * gc[1][4], w0 * gc[0][5] + w1 * gc[1][5], w0 * gc[0][6] + w1 * gc[1][6], w0 * gc[0][7] + w1 * gc[1][7]);
            } else {
                float* a1 = G.grad + off1 + 4 * t;
                red_add_v4(a0, w0 * gc[0][0], w0 * gc[0][1], w0 * gc[0][2], w0 * gc[0][3]);
                red_add_v4(a0 + 16, w0 * gc[0][4], w0 * gc[0][5], w0 * gc[0][6], w0 * gc[0][7]);
                red_add_v4(a1, w1 * gc[1][0], w1 * gc[1][1], w1 * gc[1][2], w1 * gc[1][3]);
                red_add_v4(a1 + 16, w1 * gc[1][4], w1 * gc[1][5], w1 * gc[1][6], w1 * gc[1][7]);
            }
        }
    }
}

// ---- warp-aggregated scatter (north_star item 5) ---------------------------------------------------------------------------
// The 16 samples of a tile lie on ONE ray, sorted along it (Renderer.cpp:119), 1-2 cm apart inside the near-surface band
// (Renderer.cpp:88): consecutive samples mostly share a voxel cell, and consecutive cells mostly share a face.  Instead of two to
// four vector reductions per sample and corner, the tile is transposed through a per-warp shared-memory scratch so that LANE =
// CHANNEL, and the warp walks its 16 samples serially with eight vertex sums in registers:
//     acc[j] += w_j(s) * g_c(s)[lane]
// The register slot j of a vertex is the PARITY of its absolute voxel coordinates (x&1 | (y&1)<<1 | (z&1)<<2): the eight vertices
// of a cell cover the eight parities, and a vertex shared by two consecutive cells keeps its slot -- so when the ray moves to a
// neighbouring cell only the vertices that LEAVE are flushed (one `red.global.add.f32` per vertex: 32 lanes = one whole 128-byte
// line, four full sectors) and the shared ones simply keep accumulating: one reduction per distinct vertex of the tile, whatever
// the kind of neighbour (face, edge, corner), with no register permutation.  Which slots leave is an 8-bit mask computed once per
// sample by 16 lanes in parallel; every branch of the serial walk is warp-uniform.  Measured on the bench workload: 79 -> 25
// lines per tile on the middle grid, 95 -> 42 on the fine grid.  The trilinear setup runs once per sample (lanes t = 0, 1 of a
// quad take one sample each) instead of once per lane.  Same sums as grid_sampler_3d_backward, associated per vertex.
constexpr int SCAT_GC = 40;                                  // row stride (floats) of the transposed gradient tile: conflict-free STS.128
constexpr int SCAT_W = TILE * SCAT_GC;                       // vertex weights by slot [16][8]
constexpr int SCAT_OFF = SCAT_W + TILE * 8;                  // vertex offsets (floats into the grid) by slot [16][8]
constexpr int SCAT_CELL = SCAT_OFF + TILE * 8;               // packed cell coordinates [16], then flush masks [16]
constexpr int SCAT_FLOATS = SCAT_CELL + 2 * TILE;            // per warp: 928 floats = 3712 bytes

__device__ __forceinline__ void red_add_f32(float* addr, float v) {
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr), "f"(v) : "memory");
}

// slots (parity of the coordinate along one axis, bit 0 / 1) of the old cell's two vertices along that axis that are also
// vertices of the new cell: c -> c keeps both, c -> c+1 keeps the upper one, c -> c-1 the lower one, anything else none
__device__ __forceinline__ uint32_t axis_keep(int c_old, int c_new, uint32_t even_slots, uint32_t odd_slots) {
    const int d = c_new - c_old;
    if (d == 0) return 0xffu;
    if (d == 1) return ((c_old + 1) & 1) ? odd_slots : even_slots;
    if (d == -1) return (c_old & 1) ? odd_slots : even_slots;
    return 0u;
}

// gc[r][8]: gradient of the thread's channels 4t..4t+3 and 16+4t..16+4t+3 for its samples 2g (r = 0) and 2g+1 (r = 1).
__device__ __forceinline__ void scatter_tile(const GridView& G, const Bound& bnd, const float (&p)[2][3], const float (&gc)[2][8],
                                          float* __restrict__ scr, int g, int t, int lane) {
    int* const scri = reinterpret_cast<int*>(scr);
    __syncwarp();                                            // the previous tile's reads of the scratch are done
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        float* row = scr + (2 * g + r) * SCAT_GC;
        *reinterpret_cast<float4*>(row + 4 * t) = make_float4(gc[r][0], gc[r][1], gc[r][2], gc[r][3]);
        *reinterpret_cast<float4*>(row + 16 + 4 * t) = make_float4(gc[r][4], gc[r][5], gc[r][6], gc[r][7]);
    }
    if (t < 2) {                                             // one lane per sample: lanes t = 0 / 1 of quad g take samples 2g / 2g+1
        const float q[3] = {t ? p[1][0] : p[0][0], t ? p[1][1] : p[0][1], t ? p[1][2] : p[0][2]};
        Tri s;
        tri_setup(G, bnd, q, s);
        const int smp = 2 * g + t;
        const int pc = (s.i0[0] & 1) | ((s.i0[1] & 1) << 1) | ((s.i0[2] & 1) << 2);
#pragma unroll
        for (int k = 0; k < 8; ++k) {                       // corner k of the cell -> slot k ^ pc
            int off;
            const float w = tri_corner(G, s, k, off);        // i1 is clamped at the border (weight exactly 0 there)
            scr[SCAT_W + 8 * smp + (k ^ pc)] = w;
            scri[SCAT_OFF + 8 * smp + (k ^ pc)] = off * 4;      // byte offset (unsigned 32-bit: a grid is far below 4 GB)
        }
        scri[SCAT_CELL + smp] = s.i0[0] | (s.i0[1] << 10) | (s.i0[2] << 20);
    }
    __syncwarp();
    if (lane < TILE) {                                       // slots that leave between sample lane-1 and sample lane
        uint32_t fm = 0;
        if (lane > 0) {
            const int a = scri[SCAT_CELL + lane - 1], b = scri[SCAT_CELL + lane];
            if (a != b) {
                const uint32_t keep = axis_keep(a & 1023, b & 1023, 0x55u, 0xaau) & axis_keep((a >> 10) & 1023, (b >> 10) & 1023, 0x33u, 0xccu) &
                                      axis_keep(a >> 20, b >> 20, 0x0fu, 0xf0u);
                fm = ~keep & 0xffu;
            }
        }
        scri[SCAT_CELL + TILE + lane] = (int)fm;
    }
    __syncwarp();
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.0f;
    const char* const base = reinterpret_cast<const char*>(G.grad + lane);   // + unsigned byte offset: two integer adds per address
#pragma unroll 1
    for (int s = 0; s <= TILE; ++s) {
        const uint32_t fm = s == TILE ? 0xffu : (uint32_t)scri[SCAT_CELL + TILE + s];
        if (fm) {                                            // warp-uniform: flush the vertices that leave with the previous sample's cell
            const int4 oa = *reinterpret_cast<const int4*>(scri + SCAT_OFF + 8 * (s - 1)), ob = *reinterpret_cast<const int4*>(scri + SCAT_OFF + 8 * (s - 1) + 4);
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
        if (s == TILE) break;
        const float4 wa = *reinterpret_cast<const float4*>(scr + SCAT_W + 8 * s), wb = *reinterpret_cast<const float4*>(scr + SCAT_W + 8 * s + 4);
        const float v = scr[s * SCAT_GC + lane];
        acc[0] = fmaf(wa.x, v, acc[0]); acc[1] = fmaf(wa.y, v, acc[1]); acc[2] = fmaf(wa.z, v, acc[2]); acc[3] = fmaf(wa.w, v, acc[3]);
        acc[4] = fmaf(wb.x, v, acc[4]); acc[5] = fmaf(wb.y, v, acc[5]); acc[6] = fmaf(wb.z, v, acc[6]); acc[7] = fmaf(wb.w, v, acc[7]);
    }
}

// Backward of one decoder for one tile.  gout[r][o] = cotangent of the decoder outputs of the thread's two rows.
// On return: gcf[r][8] = d L / d c for the thread's 8 grid channels (first 32 channels only: the middle half of
// the fine decoder's input is stop-gradient, MLP.cpp:79-84); gp[r][3] = partial d L / d p from the Fourier path.
template <int C, int O, bool P3, bool NEED_C, bool NEED_E, bool STASH>
__device__ __forceinline__ void decoder_backward(const float* __restrict__ sm, const float (&p)[2][3], int g, int t,
                                                 const float (&gout)[2][4], const uint32_t (&masks)[5],
                                                 float (&gcf)[2][8], float (&gp)[2][3], float* st0, float* st1) {
    using L = DecSmem<C>;
    constexpr int NO = O == 4 ? 3 : 1;
    float gh[4][4], gu[4][4], gc[4][4], gu0[4][4], gu3[4][4];
    // output layer: g_h5 = g_out Wo
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        gh[j][0] = gh[j][1] = gh[j][2] = gh[j][3] = 0.0f;
#pragma unroll
        for (int o = 0; o < NO; ++o) {
            const float2 w = *reinterpret_cast<const float2*>(sm + L::WO + o * HID + 8 * j + 2 * t);
            gh[j][0] = fmaf(gout[0][o], w.x, gh[j][0]); gh[j][1] = fmaf(gout[0][o], w.y, gh[j][1]);
            gh[j][2] = fmaf(gout[1][o], w.x, gh[j][2]); gh[j][3] = fmaf(gout[1][o], w.y, gh[j][3]);
        }
    }
    if (NEED_C) {   // g_c starts with the output layer's share g_out (Wo Fc_4); accumulator column 8j+2t+b <-> channel fc_channel_bwd(.)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int b = 0; b < 2; ++b) {
                const int ch = fc_channel_bwd(8 * j + 2 * t + b);
                float s0 = 0.0f, s1 = 0.0f;
#pragma unroll
                for (int o = 0; o < NO; ++o) { const float w = sm[L::WOC + o * C + ch]; s0 = fmaf(gout[0][o], w, s0); s1 = fmaf(gout[1][o], w, s1); }
                gc[j][b] = s0; gc[j][2 + b] = s1;
            }
    }
#pragma unroll
    for (int i = 4; i >= 0; --i) {
        apply_mask(gu, gh, masks[i]);
        if (STASH) stash_tile(st0, st1, stash::GU + HID * i, gu, t);
        if (NEED_E && i == 3) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int q = 0; q < 4; ++q) gu3[j][q] = gu[j][q];
        }
        if (i == 0) {
            if (NEED_E) {
#pragma unroll
                for (int j = 0; j < 4; ++j)
#pragma unroll
                    for (int q = 0; q < 4; ++q) gu0[j][q] = gu[j][q];
            }
        } else {   // g_h_i = g_u W_i, and with the same A fragments g_c += g_u G_{i-1} (composed grid-feature path)
            zero_tile(gh);
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) {
                AFrag<P3> a;
                afrag_from_c<P3>(a, gu[2 * kk], gu[2 * kk + 1]);
                kstep_fwd<P3, 4>(gh, a, wmat(sm, L::w(i)), L::SH, kk, g, t);                // W_i^T
                if (NEED_C) kstep_fwd<P3, 4>(gc, a, wmat(sm, L::FC + (i - 1) * HID * L::SH), L::SH, kk, g, t);   // G_{i-1}^T [channel][out]
            }
        }
    }
    if (NEED_C) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            gcf[0][2 * j] = gc[j][0]; gcf[0][2 * j + 1] = gc[j][1];
            gcf[1][2 * j] = gc[j][2]; gcf[1][2 * j + 1] = gc[j][3];
        }
    }
    if (NEED_E) {
        // g_e = g_u0 W0 + g_u3 W3e, four 8-feature tiles (four independent accumulator chains) at a time; then the chain
        // through e = sin(p B)
        AFrag<P3> a0[2], a3[2];
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) { afrag_from_c<P3>(a0[kk], gu0[2 * kk], gu0[2 * kk + 1]); afrag_from_c<P3>(a3[kk], gu3[2 * kk], gu3[2 * kk + 1]); }
#pragma unroll 1
        for (int jg = 0; jg < EMBP / 32; ++jg) {
            float ge[4][4];
            zero_tile(ge);
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) {
                kstep_fwd<P3, 4>(ge, a0[kk], wmat(sm, L::W0 + 32 * jg * L::SH), L::SH, kk, g, t);    // W0^T rows 32 jg ..
                kstep_fwd<P3, 4>(ge, a3[kk], wmat(sm, L::W3E + 32 * jg * L::SH), L::SH, kk, g, t);
            }
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                const int f0 = 32 * jg + 8 * jj + 2 * t;
                const float2 B0 = *reinterpret_cast<const float2*>(sm + L::B + f0);
                const float2 B1 = *reinterpret_cast<const float2*>(sm + L::B + EMBP + f0);
                const float2 B2 = *reinterpret_cast<const float2*>(sm + L::B + 2 * EMBP + f0);
                float sn, c00, c01, c10, c11;
                ff_sincos(fmaf(p[0][2], B2.x, fmaf(p[0][1], B1.x, p[0][0] * B0.x)), sn, c00);
                ff_sincos(fmaf(p[0][2], B2.y, fmaf(p[0][1], B1.y, p[0][0] * B0.y)), sn, c01);
                ff_sincos(fmaf(p[1][2], B2.x, fmaf(p[1][1], B1.x, p[1][0] * B0.x)), sn, c10);
                ff_sincos(fmaf(p[1][2], B2.y, fmaf(p[1][1], B1.y, p[1][0] * B0.y)), sn, c11);
                const float q00 = ge[jj][0] * c00, q01 = ge[jj][1] * c01, q10 = ge[jj][2] * c10, q11 = ge[jj][3] * c11;
                if (STASH) {
                    __stcs(reinterpret_cast<float2*>(st0 + stash::GE + f0), make_float2(q00, q01));
                    __stcs(reinterpret_cast<float2*>(st1 + stash::GE + f0), make_float2(q10, q11));
                }
                gp[0][0] += q00 * B0.x + q01 * B0.y; gp[0][1] += q00 * B1.x + q01 * B1.y; gp[0][2] += q00 * B2.x + q01 * B2.y;
                gp[1][0] += q10 * B0.x + q11 * B0.y; gp[1][1] += q10 * B1.x + q11 * B1.y; gp[1][2] += q10 * B2.x + q11 * B2.y;
            }
        }
    }
}

template <int C, int O, bool P3, bool GRID, bool RAY, bool WG>
__device__ __forceinline__ void backward_tile(const DecodeParams& P, const float* __restrict__ sm, float* __restrict__ scr, int dec, int base,
                                              int g, int t, int lane) {
    float p[2][3]; int sidx[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) sidx[r] = base + 2 * g + r;
    const int ray = base / P.S;
    // every global input of the tile is requested before the first use (one exposed memory latency instead of a chain of four)
    const uint8_t ok = P.valid ? P.valid[ray] : (uint8_t)1;
    const float o[3] = {P.rays_o[3 * ray], P.rays_o[3 * ray + 1], P.rays_o[3 * ray + 2]};
    const float d[3] = {P.rays_d[3 * ray], P.rays_d[3 * ray + 1], P.rays_d[3 * ray + 2]};
    const float zz[2] = {P.z[sidx[0]], P.z[sidx[1]]};
    const float4 gr[2] = {*reinterpret_cast<const float4*>(P.g_raw + 4 * (size_t)sidx[0]), *reinterpret_cast<const float4*>(P.g_raw + 4 * (size_t)sidx[1])};
    // relu masks saved by the training forward (k_decode_fwd<.., TRAIN>): the data gradient needs nothing else
    uint32_t mw[10];
    const bool per_sample = (P.mask_layout >> dec) & 1;   // bit d: decoder d's masks are one word per sample and layer (tcgen05 forward)
    if (!per_sample) {
        const uint32_t* mb = P.masks + ((size_t)(dec - 1) * (P.P / TILE) + base / TILE) * 96 + lane;
        mw[0] = mb[0]; mw[1] = mb[32]; mw[2] = mb[64];
    } else {   // one word per sample and layer (bit f = relu of feature f)
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            const uint32_t* mb = P.masks + ((size_t)(dec - 1) * 5 + i) * P.mask_stride;
            mw[2 * i] = mb[sidx[0]]; mw[2 * i + 1] = mb[sidx[1]];
        }
    }
    if (!ok) return;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
#pragma unroll
        for (int a = 0; a < 3; ++a) p[r][a] = __fadd_rn(o[a], __fmul_rn(d[a], zz[r]));
    }
    float gout[2][4];
    bool any = false;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        if (O == 4) { gout[r][0] = gr[r].x; gout[r][1] = gr[r].y; gout[r][2] = gr[r].z; gout[r][3] = 0.0f; any = any || gr[r].x != 0.0f || gr[r].y != 0.0f || gr[r].z != 0.0f; }
        else { gout[r][0] = gr[r].w; gout[r][1] = gout[r][2] = gout[r][3] = 0.0f; any = any || gr[r].w != 0.0f; }
    }
    if (!WG && !__any_sync(0xffffffffu, any)) return;   // nothing flows into this tile

    uint32_t masks[5];
    if (!per_sample) {
        masks[0] = mw[0] & 0xffffu; masks[1] = mw[0] >> 16; masks[2] = mw[1] & 0xffffu; masks[3] = mw[1] >> 16; masks[4] = mw[2];
    } else {   // pick this thread's fragment bits out of the per-sample words
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            const uint32_t w0 = mw[2 * i] >> (2 * t), w1 = mw[2 * i + 1] >> (2 * t);
            uint32_t m = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) m |= (((w0 >> (8 * j)) & 3u) << (4 * j)) | (((w1 >> (8 * j)) & 3u) << (4 * j + 2));
            masks[i] = m;
        }
    }
    float* st0 = nullptr; float* st1 = nullptr;
    if (WG) {   // E / H / Cc columns of the stash rows were written by the forward; add the gradient side
        st0 = P.stash + (size_t)sidx[0] * stash::W; st1 = P.stash + (size_t)sidx[1] * stash::W;
        if (t == 0) {
            __stcs(reinterpret_cast<float4*>(st0 + stash::GO), make_float4(gout[0][0], gout[0][1], gout[0][2], 0.0f));
            __stcs(reinterpret_cast<float4*>(st1 + stash::GO), make_float4(gout[1][0], gout[1][1], gout[1][2], 0.0f));
            __stcs(reinterpret_cast<float4*>(st0 + stash::Pp), make_float4(p[0][0], p[0][1], p[0][2], 1.0f));
            __stcs(reinterpret_cast<float4*>(st1 + stash::Pp), make_float4(p[1][0], p[1][1], p[1][2], 1.0f));
        }
    }
    float gcf[2][8], gp[2][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};
    decoder_backward<C, O, P3, GRID || RAY, RAY || WG, WG>(sm, p, g, t, gout, masks, gcf, gp, st0, st1);
#if NSB_SCATTER_AGG
    if (RAY) grid_backward<false, true>(P.grid[dec], P.bnd, p, gcf, t, gp);
    if (GRID) scatter_tile(P.grid[dec], P.bnd, p, gcf, scr, g, t, lane);
#else
    if (GRID || RAY) grid_backward<GRID, RAY>(P.grid[dec], P.bnd, p, gcf, t, gp);
#endif
    if (RAY) {
        // gp holds per-thread partials (over the quad's channels / features): finish the quad sum, then the
        // 16 rows of the tile belong to one ray: d L/d o = sum g_p, d L/d d = sum z g_p
        float acc[6];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const float g0 = quad_sum(gp[0][a]), g1 = quad_sum(gp[1][a]);
            acc[a] = g0 + g1;
            acc[3 + a] = zz[0] * g0 + zz[1] * g1;
        }
#pragma unroll
        for (int a = 0; a < 6; ++a) {
            float v = acc[a];
            v += __shfl_xor_sync(0xffffffffu, v, 4); v += __shfl_xor_sync(0xffffffffu, v, 8); v += __shfl_xor_sync(0xffffffffu, v, 16);
            if (lane == 0) atomicAdd(P.d_rays + 6 * (size_t)ray + a, v);
        }
    }
}

template <bool P3, bool GRID, bool RAY, bool WG>
__global__ void __launch_bounds__(BWD_THREADS, NSB_BWD_MIN_CTAS) k_decode_bwd(const DecodeParams P) {
    extern __shared__ __align__(128) float sm[];
    int dec = 1;
#pragma unroll
    for (int d = 2; d < 4; ++d) if ((int)blockIdx.x >= P.cta_begin[d]) dec = d;
    const int cta = blockIdx.x - P.cta_begin[dec], ncta = P.cta_begin[dec + 1] - P.cta_begin[dec];
    if (dec == 2) load_decoder_image<64>(sm, P.wimg_bwd[2], threadIdx.x, blockDim.x);
    else load_decoder_image<32>(sm, P.wimg_bwd[dec], threadIdx.x, blockDim.x);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int ntiles = P.P / TILE;
    (void)cta; (void)ncta;
    float* scr = sm + DecSmem<64>::TOTAL + warp * SCAT_FLOATS;     // per-warp scratch of the aggregated scatter, behind the largest image
    TileQueue q; q.init(P.tile_ctr + dec, 0ull, ntiles, lane);
    for (int tile = q.next(lane); tile >= 0; tile = q.next(lane)) {
        if (dec == 1) backward_tile<32, 1, P3, GRID, RAY, false>(P, sm, scr, 1, tile * TILE, g, t, lane);
        else if (dec == 2) backward_tile<64, 1, P3, GRID, RAY, false>(P, sm, scr, 2, tile * TILE, g, t, lane);
        else backward_tile<32, 4, P3, GRID, RAY, WG>(P, sm, scr, 3, tile * TILE, g, t, lane);
    }
}

}  // namespace nsb
