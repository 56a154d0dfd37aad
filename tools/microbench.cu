// Microbenchmarks that calibrate the roofline denominators MEASURED_PEAKS.json does not carry:
//   * mma.sync m16n8k8 tf32 issue rate (the legacy warp-level tensor path the decoders use),
//   * L2-resident gather bandwidth for 128-byte voxel lines (the grids are 11.2 MB, far below the 126 MB L2),
//   * fp32 vector-reduction (red.global.add.v4.f32) throughput into an L2-resident buffer.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench tools/microbench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>

__global__ void k_mma(float* out, int iters) {
    float d[8][4];
    for (int i = 0; i < 8; ++i) d[i][0] = d[i][1] = d[i][2] = d[i][3] = 0.f;
    uint32_t a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, b0 = a0 * 7, b1 = a0 * 5;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(d[i][0]), "+f"(d[i][1]), "+f"(d[i][2]), "+f"(d[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
    float s = 0; for (int i = 0; i < 8; ++i) s += d[i][0] + d[i][1] + d[i][2] + d[i][3];
    if (s == 12345.f) out[0] = s;
}

__global__ void k_mma_bf16(float* out, int iters) {
    float d[8][4];
    for (int i = 0; i < 8; ++i) d[i][0] = d[i][1] = d[i][2] = d[i][3] = 0.f;
    uint32_t a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, b0 = a0 * 7, b1 = a0 * 5;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(d[i][0]), "+f"(d[i][1]), "+f"(d[i][2]), "+f"(d[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
    float s = 0; for (int i = 0; i < 8; ++i) s += d[i][0] + d[i][1] + d[i][2] + d[i][3];
    if (s == 12345.f) out[0] = s;
}

// mixed stream as the hybrid split would issue it: one tf32 m16n8k8 + one bf16 m16n8k16 per accumulator step
__global__ void k_mma_mix(float* out, int iters) {
    float d[8][4];
    for (int i = 0; i < 8; ++i) d[i][0] = d[i][1] = d[i][2] = d[i][3] = 0.f;
    uint32_t a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, b0 = a0 * 7, b1 = a0 * 5;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(d[i][0]), "+f"(d[i][1]), "+f"(d[i][2]), "+f"(d[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(d[i][0]), "+f"(d[i][1]), "+f"(d[i][2]), "+f"(d[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
        }
    }
    float s = 0; for (int i = 0; i < 8; ++i) s += d[i][0] + d[i][1] + d[i][2] + d[i][3];
    if (s == 12345.f) out[0] = s;
}

__global__ void k_gather(const float4* __restrict__ buf, uint32_t nlines, float* out, int iters) {
    // each quad (4 lanes x 2 float4) reads one random 128-byte line per step, like the trilinear corner fetch
    uint32_t quad = (blockIdx.x * blockDim.x + threadIdx.x) >> 2, t = threadIdx.x & 3;
    uint32_t s = quad * 2654435761u + 12345u;
    float acc = 0.f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            s = s * 1664525u + 1013904223u;
            const uint32_t line = (s >> 8) % nlines;
            const float4 v0 = __ldg(buf + line * 8 + 2 * t), v1 = __ldg(buf + line * 8 + 2 * t + 1);
            acc += v0.x + v0.w + v1.y + v1.z;
        }
    }
    if (acc == 12345.f) out[0] = acc;
}

__global__ void k_red(float* buf, uint32_t nlines, int iters) {
    uint32_t quad = (blockIdx.x * blockDim.x + threadIdx.x) >> 2, t = threadIdx.x & 3;
    uint32_t s = quad * 2654435761u + 777u;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            s = s * 1664525u + 1013904223u;
            const uint32_t line = (s >> 8) % nlines;
            float* a = buf + line * 32 + 8 * t;
            asm volatile("red.global.add.v4.f32 [%0], {%1,%1,%1,%1};" ::"l"(a), "f"(1.0f) : "memory");
            asm volatile("red.global.add.v4.f32 [%0], {%1,%1,%1,%1};" ::"l"(a + 4), "f"(1.0f) : "memory");
        }
    }
}

// variants of the scatter pattern: MODE 0 = lane t adds floats 8t..8t+7 (two v4: each touches half of four sectors),
// MODE 1 = lane t adds floats 4t..4t+3 then 16+4t..16+4t+3 (each v4 instruction of a quad fills two whole sectors);
// SAME = all eight quads of a warp hit the same line (neighbouring samples of a ray in one voxel cell)
template <int MODE, bool SAME>
__global__ void k_red2(float* buf, uint32_t nlines, int iters) {
    uint32_t quad = (blockIdx.x * blockDim.x + threadIdx.x) >> (SAME ? 5 : 2), t = threadIdx.x & 3;
    uint32_t s = quad * 2654435761u + 777u;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            s = s * 1664525u + 1013904223u;
            const uint32_t line = (s >> 8) % nlines;
            float* a = buf + line * 32 + (MODE == 0 ? 8 * t : 4 * t);
            asm volatile("red.global.add.v4.f32 [%0], {%1,%1,%1,%1};" ::"l"(a), "f"(1.0f) : "memory");
            asm volatile("red.global.add.v4.f32 [%0], {%1,%1,%1,%1};" ::"l"(a + (MODE == 0 ? 4 : 16)), "f"(1.0f) : "memory");
        }
    }
}

// Does legacy HMMA work overlap with FP32 ALU work?  MODE 0: every warp issues 8 independent HMMAs + NALU independent FFMAs per
// step; MODE 1: even warps only HMMAs (16 per step), odd warps only FFMAs (2 NALU per step); MODE 2: FFMAs only; MODE 3: HMMAs only.
template <int MODE, int NALU>
__global__ void k_mma_alu(float* out, int iters) {
    float d[8][4], f[16];
    for (int i = 0; i < 8; ++i) d[i][0] = d[i][1] = d[i][2] = d[i][3] = 0.f;
    for (int i = 0; i < 16; ++i) f[i] = threadIdx.x * 0.001f + i;
    uint32_t a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, b0 = a0 * 7, b1 = a0 * 5;
    const bool odd = (threadIdx.x >> 7) & 1;   // warps 4..7, 12..15: every scheduler (warp % 4) gets two warps of each kind
    const float m = 1.0001f, c = 0.5f;
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0 || MODE == 3 || (MODE == 1 && !odd)) {
#pragma unroll
            for (int rep = 0; rep < (MODE == 1 ? 2 : 1); ++rep)
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                                 : "+f"(d[i][0]), "+f"(d[i][1]), "+f"(d[i][2]), "+f"(d[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
        }
        if (MODE == 0 || MODE == 2 || (MODE == 1 && odd)) {
#pragma unroll
            for (int rep = 0; rep < (MODE == 1 ? 2 : 1) * NALU / 16; ++rep)
#pragma unroll
                for (int i = 0; i < 16; ++i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(m), "f"(c));
        }
    }
    float s = 0; for (int i = 0; i < 8; ++i) s += d[i][0] + d[i][1] + d[i][2] + d[i][3];
    for (int i = 0; i < 16; ++i) s += f[i];
    if (s == 12345.f) out[0] = s;
}
template <int MODE, int NALU>
static void run_mma_alu(float* out, int sms, int clk, const char* name, cudaEvent_t e0, cudaEvent_t e1) {
    const int iters = 4000, warps = 16; float ms;
    k_mma_alu<MODE, NALU><<<sms, warps * 32>>>(out, 50);
    cudaEventRecord(e0); k_mma_alu<MODE, NALU><<<sms, warps * 32>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    printf(", \"%s_clk_per_step\": %.1f", name, ms * 1e-3 * clk * 1e6 / iters);   // SM clocks per loop step (16 warps)
}

// whole-line reductions: one warp instruction adds 32 consecutive floats (lane = channel) of a random 128-byte line
__global__ void k_red_line(float* buf, uint32_t nlines, int iters) {
    uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    uint32_t s = warp * 2654435761u + 777u;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            s = s * 1664525u + 1013904223u;
            const uint32_t line = (s >> 8) % nlines;
            asm volatile("red.global.add.f32 [%0], %1;" ::"l"(buf + line * 32 + lane), "f"(1.0f) : "memory");
        }
    }
}
template <int MODE, bool SAME>
static void run_red2(float* buf, uint32_t nlines, int sms, const char* name, cudaEvent_t e0, cudaEvent_t e1) {
    const int grid = sms * 8, iters = 50; float ms;
    k_red2<MODE, SAME><<<grid, 256>>>(buf, nlines, 5);
    cudaEventRecord(e0); k_red2<MODE, SAME><<<grid, 256>>>(buf, nlines, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    const double bytes = 128.0 * 8 * iters * (double)grid * 256 / 4;
    printf(", \"%s\": %.0f", name, bytes / ms * 1e-6);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_mhz\": %d", p.name, p.multiProcessorCount, clk / 1000);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float* out; cudaMalloc(&out, 64);
    float ms;
    for (int warps : {4, 8, 16, 32}) {
        const int iters = 20000, grid = p.multiProcessorCount;
        k_mma<<<grid, warps * 32>>>(out, 100);
        cudaEventRecord(e0); k_mma<<<grid, warps * 32>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        const double flop = 2.0 * 16 * 8 * 8 * 8.0 * iters * warps * grid;
        printf(", \"mma_tf32_tflops_w%d\": %.1f", warps, flop / ms * 1e-9);
    }
    {
        const int iters = 20000, grid = p.multiProcessorCount, warps = 16;
        k_mma_bf16<<<grid, warps * 32>>>(out, 100);
        cudaEventRecord(e0); k_mma_bf16<<<grid, warps * 32>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        printf(", \"mma_f16_k16_instr_per_clk_sm\": %.3f, \"mma_f16_tflops\": %.1f", 8.0 * iters * warps / (ms * 1e-3 * clk * 1e3), 2.0 * 16 * 8 * 16 * 8.0 * iters * warps * grid / ms * 1e-9);
        k_mma<<<grid, warps * 32>>>(out, 100);
        cudaEventRecord(e0); k_mma<<<grid, warps * 32>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        printf(", \"mma_tf32_k8_instr_per_clk_sm\": %.3f", 8.0 * iters * warps / (ms * 1e-3 * clk * 1e3));
        k_mma_mix<<<grid, warps * 32>>>(out, 100);
        cudaEventRecord(e0); k_mma_mix<<<grid, warps * 32>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        printf(", \"mma_mix_pairs_per_clk_sm\": %.3f", 8.0 * iters * warps / (ms * 1e-3 * clk * 1e3));
    }
    const uint32_t nlines = 11 * 1024 * 1024 / 128;
    float4* buf; cudaMalloc(&buf, (size_t)nlines * 128); cudaMemset(buf, 0, (size_t)nlines * 128);
    for (int occ : {2, 4, 8}) {
        const int grid = p.multiProcessorCount * occ, iters = 200;
        k_gather<<<grid, 256>>>(buf, nlines, out, 10);
        cudaEventRecord(e0); k_gather<<<grid, 256>>>(buf, nlines, out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        const double bytes = 128.0 * 8 * iters * (double)grid * 256 / 4;
        printf(", \"l2_gather_gbs_occ%d\": %.0f", occ, bytes / ms * 1e-6);
    }
    for (int occ : {2, 8}) {
        const int grid = p.multiProcessorCount * occ, iters = 50;
        k_red<<<grid, 256>>>((float*)buf, nlines, 5);
        cudaEventRecord(e0); k_red<<<grid, 256>>>((float*)buf, nlines, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        const double bytes = 128.0 * 8 * iters * (double)grid * 256 / 4;
        printf(", \"l2_redv4_gbs_occ%d\": %.0f", occ, bytes / ms * 1e-6);
    }
    run_red2<0, false>((float*)buf, nlines, p.multiProcessorCount, "red_mode0_gbs", e0, e1);
    run_red2<1, false>((float*)buf, nlines, p.multiProcessorCount, "red_mode1_gbs", e0, e1);
    run_red2<0, true>((float*)buf, nlines, p.multiProcessorCount, "red_mode0_sameline_gbs", e0, e1);
    run_red2<1, true>((float*)buf, nlines, p.multiProcessorCount, "red_mode1_sameline_gbs", e0, e1);

    {
        const int grid = p.multiProcessorCount * 8, iters = 200;
        k_red_line<<<grid, 256>>>((float*)buf, nlines, 5);
        cudaEventRecord(e0); k_red_line<<<grid, 256>>>((float*)buf, nlines, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        printf(", \"red_line_f32_gbs\": %.0f", 128.0 * 8 * iters * (double)grid * 8 / ms * 1e-6);
    }
    run_mma_alu<3, 64>(out, p.multiProcessorCount, clk / 1000, "hmma8_only", e0, e1);
    run_mma_alu<2, 64>(out, p.multiProcessorCount, clk / 1000, "ffma64_only", e0, e1);
    run_mma_alu<0, 64>(out, p.multiProcessorCount, clk / 1000, "hmma8_ffma64_same_warp", e0, e1);
    run_mma_alu<1, 64>(out, p.multiProcessorCount, clk / 1000, "hmma16_ffma128_alternate_warps", e0, e1);
    run_mma_alu<0, 128>(out, p.multiProcessorCount, clk / 1000, "hmma8_ffma128_same_warp", e0, e1);
    run_mma_alu<2, 128>(out, p.multiProcessorCount, clk / 1000, "ffma128_only", e0, e1);
    printf(", \"err\": \"%s\"}\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
