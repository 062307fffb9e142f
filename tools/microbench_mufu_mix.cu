// MUFU.EX2 throughput next to the other instructions of the attention softmax loop (B200).
// tools/microbench.cu measured 21.3 ex2/clk/SM in a pure loop but the softmax replica sits at 15.6: which co-issued
// instruction class costs the difference, and does issuing the exponentials in bursts recover it?
// Every thread runs ITER iterations over 64 independent values; variants differ in what accompanies the 64 MUFUs.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/mb_mufu tools/microbench_mufu_mix.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t pack2(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void unpack2(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) { uint64_t d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t cvt2(float lo, float hi) { uint32_t r; asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo)); return r; }

// VAR bit 0: FFMA2 producing the arguments; bit 1: FADD2 row sums; bit 2: F2FP packs; bit 3: feed the result back (dependent
// chain across iterations through the value itself instead of fresh arguments)
template <int VAR, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) k(const float* __restrict__ in, float* out, long long* cycles, int iters, float c, float m) {
    float s[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) s[i] = in[(threadIdx.x * 64 + i) % 4096];
    uint64_t acc0 = 0, acc1 = 0;
    uint32_t x = 0;
    float mneg = m;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
        const uint64_t C2 = pack2(c, c), M2 = pack2(mneg, mneg);
        mneg += 1e-6f;
#pragma unroll
        for (int i = 0; i < 64; i += 2) {
            float x0 = s[i], x1 = s[i + 1];
            if (VAR & 1) unpack2(ffma2(pack2(x0, x1), C2, M2), x0, x1);
            const float p0 = ex2(x0), p1 = ex2(x1);
            if (VAR & 2) { if (i & 2) acc1 = fadd2(acc1, pack2(p0, p1)); else acc0 = fadd2(acc0, pack2(p0, p1)); }
            if (VAR & 4) x ^= cvt2(p0, p1);
            if (!(VAR & 1)) { s[i] = p0; s[i + 1] = p1; }   // pure chain: ex2 of ex2 (stays finite: 2^x of values in (0, 1])
            else if (!(VAR & 6)) { s[i] += p0 * 1e-30f; s[i + 1] += p1 * 1e-30f; }   // keep the results alive cheaply
        }
    }
    const long long t1 = clock64();
    float a, b, cc, d;
    unpack2(acc0, a, b); unpack2(acc1, cc, d);
    float r = a + b + cc + d + __uint_as_float(x) + mneg;
#pragma unroll
    for (int i = 0; i < 64; ++i) r += s[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int VAR, int WARPS>
void run(const char* name) {
    const int blocks = 148, iters = 4000;
    float *in, *out; long long* cyc;
    cudaMalloc(&in, 4096 * 4); cudaMalloc(&out, blocks * WARPS * 32 * 4); cudaMalloc(&cyc, blocks * 8);
    float h[4096];
    for (int i = 0; i < 4096; ++i) h[i] = -0.01f * (i % 977);
    cudaMemcpy(in, h, sizeof h, cudaMemcpyHostToDevice);
    k<VAR, WARPS><<<blocks, WARPS * 32>>>(in, out, cyc, 10, 1.0f, 0.0f);
    k<VAR, WARPS><<<blocks, WARPS * 32>>>(in, out, cyc, iters, 1.0f, 0.0f);
    cudaDeviceSynchronize();
    long long hc[148];
    cudaMemcpy(hc, cyc, sizeof hc, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < blocks; ++i) avg += hc[i]; avg /= blocks;
    const double per_clk = double(iters) * 64 * WARPS * 32 / avg;
    printf("{\"variant\": \"%s\", \"warps\": %d, \"ex2_per_clk_per_sm\": %.2f, \"cycles_per_warp_mufu\": %.2f, \"err\": \"%s\"}\n", name, WARPS, per_clk,
           32.0 * 4 / per_clk, cudaGetErrorString(cudaGetLastError()));
    cudaFree(in); cudaFree(out); cudaFree(cyc);
}

int main() {
    run<0, 4>("ex2 chain only");
    run<7, 4>("ffma2 + ex2 + fadd2 + f2fp");
    run<0, 8>("ex2 chain only");
    run<0, 16>("ex2 chain only");
    run<1, 8>("ffma2 + ex2");
    run<3, 8>("ffma2 + ex2 + fadd2");
    run<5, 8>("ffma2 + ex2 + f2fp");
    run<7, 8>("ffma2 + ex2 + fadd2 + f2fp");
    run<7, 16>("ffma2 + ex2 + fadd2 + f2fp");
    run<6, 8>("ex2 chain + fadd2 + f2fp");
    run<2, 8>("ex2 chain + fadd2");
    run<4, 8>("ex2 chain + f2fp");
    return 0;
}
