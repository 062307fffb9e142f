// Throughput microbenchmarks that decide the softmax design of the attention kernel on B200:
//   MUFU.EX2 fp32 vs packed f16x2 / bf16x2, FFMA vs packed FFMA2 (fma.rn.f32x2), FMNMX 2- vs 3-input.
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench tools/microbench.cu
// Prints results per op in (lane-results / clk / SM).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 4096
#define UNROLL 8

template <int OP>
__global__ void __launch_bounds__(1024) bench(float* out, long long* cycles, float seed) {
    float a[UNROLL];
    uint32_t u[UNROLL];
    unsigned long long d[UNROLL];
#pragma unroll
    for (int i = 0; i < UNROLL; ++i) {
        a[i] = seed * (threadIdx.x + i) * 1e-3f - 1.0f;
        u[i] = 0x3c003c00u + threadIdx.x + i;   // f16x2 / bf16x2 payloads near 1.0
        d[i] = ((unsigned long long)__float_as_uint(a[i]) << 32) | __float_as_uint(a[i] * 0.5f);
    }
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < UNROLL; ++i) {
            if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
            if (OP == 1) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(u[i]));
            if (OP == 2) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(u[i]));
            if (OP == 3) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(0.999f), "f"(0.001f));
            if (OP == 4) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(d[i]) : "l"(0x3f7fbe773f7fbe77ull), "l"(0x3a83126f3a83126full));
            if (OP == 5) asm volatile("max.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(a[(i + 1) % UNROLL]));
            if (OP == 6) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(a[(i + 1) % UNROLL]), "f"(a[(i + 2) % UNROLL]));
            if (OP == 7) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a[i]));
            if (OP == 8) asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(a[i]), "f"(a[(i + 1) % UNROLL]));
            if (OP == 9) asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(a[i]), "f"(a[(i + 1) % UNROLL]));
            if (OP == 10) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(d[i]) : "l"(0x3a83126f3a83126full));
            if (OP == 11) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(d[i]) : "l"(0x3f7fbe773f7fbe77ull));
        }
    }
    long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < UNROLL; ++i) s += a[i] + __uint_as_float(u[i]) + __uint_as_float((uint32_t)d[i]) + __uint_as_float((uint32_t)(d[i] >> 32));
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char* name, int results_per_instr) {
    float* out;
    long long* cyc;
    int blocks = 148 * 2, threads = 1024;
    cudaMalloc(&out, blocks * threads * sizeof(float));
    cudaMalloc(&cyc, blocks * sizeof(long long));
    bench<OP><<<blocks, threads>>>(out, cyc, 1.0f);
    cudaDeviceSynchronize();
    bench<OP><<<blocks, threads>>>(out, cyc, 1.0f);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[296];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < blocks; ++i) avg += h[i];
    avg /= blocks;
    // 2 blocks of 1024 threads are co-resident per SM: lane-instr per SM = 2048 * ITERS * UNROLL over `avg` cycles
    double per_clk = 2048.0 * ITERS * UNROLL / avg;
    printf("{\"op\": \"%s\", \"lane_instr_per_clk_per_sm\": %.2f, \"results_per_clk_per_sm\": %.2f, \"err\": \"%s\"}\n", name, per_clk,
           per_clk * results_per_instr, cudaGetErrorString(e));
    cudaFree(out);
    cudaFree(cyc);
}

int main() {
    run<0>("ex2.approx.ftz.f32", 1);
    run<1>("ex2.approx.f16x2", 2);
    run<2>("ex2.approx.ftz.bf16x2", 2);
    run<3>("fma.rn.f32", 1);
    run<4>("fma.rn.f32x2", 2);
    run<5>("max.f32 (2-input)", 1);
    run<6>("max.f32 (3-input)", 2);
    run<7>("tanh.approx.f32", 1);
    run<8>("cvt.rn.bf16x2.f32", 2);
    run<9>("cvt.rn.f16x2.f32", 2);
    run<10>("add.rn.f32x2", 2);
    run<11>("mul.rn.f32x2", 2);
    return 0;
}
