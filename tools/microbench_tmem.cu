// TMEM load/store throughput on B200: is reading the fp32 score tile (128 x 128 x 4 B per Q-tile x KV-tile) a
// bottleneck for head_dim-64 attention (MMA budget 512 cycles per tile)?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/mb_tmem tools/microbench_tmem.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

#define LD32(taddr, r)                                                                                                   \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                               \
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"                                               \
                 "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"                              \
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),       \
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), \
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]),            \
                   "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]),            \
                   "=r"(r[30]), "=r"(r[31])                                                                              \
                 : "r"(taddr)                                                                                            \
                 : "memory")
#define ST32(taddr, r)                                                                                                   \
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "                                                         \
                 "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"                                              \
                 "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),                       \
                 "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),      \
                 "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]),          \
                 "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]),         \
                 "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])                      \
                 : "memory")

// MODE 0: loads only; MODE 1: stores only; MODE 2: 4 loads + 2 stores per "tile" (the softmax traffic pattern)
template <int MODE>
__global__ void __launch_bounds__(256, 1) tmem_bench(float* out, long long* cycles, int iters) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = slot + (uint32_t((warp & 3) * 32) << 16) + (warp >> 2) * 256;
    uint32_t r[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) r[i] = threadIdx.x + i;
    ST32(base, r); ST32(base + 32, r); ST32(base + 64, r); ST32(base + 96, r);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    uint32_t acc = 0;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0 || MODE == 2) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                uint32_t v[32];
                LD32(base + c * 32, v);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int i = 0; i < 32; ++i) acc ^= v[i];
            }
        }
        if (MODE == 1 || MODE == 2) {
            r[0] = acc + it;
            ST32(base, r);
            ST32(base + 32, r);
            if (MODE == 1) { ST32(base + 64, r); ST32(base + 96, r); }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512) : "memory");
}

template <int MODE>
void run(const char* name, int warps) {
    const int blocks = 148, iters = 4000;
    float* out;
    long long* cyc;
    cudaMalloc(&out, blocks * 256 * sizeof(float));
    cudaMalloc(&cyc, blocks * sizeof(long long));
    tmem_bench<MODE><<<blocks, warps * 32>>>(out, cyc, iters);
    cudaDeviceSynchronize();
    tmem_bench<MODE><<<blocks, warps * 32>>>(out, cyc, iters);
    cudaError_t e = cudaDeviceSynchronize();
    long long hc[148];
    cudaMemcpy(hc, cyc, sizeof(hc), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < blocks; ++i) avg += hc[i];
    avg /= blocks;
    const double bytes_ld = (MODE == 1 ? 0.0 : 4.0 * 32 * 32 * 4) * warps, bytes_st = (MODE == 0 ? 0.0 : (MODE == 1 ? 4.0 : 2.0) * 32 * 32 * 4) * warps;
    printf("{\"variant\": \"%s\", \"warps\": %d, \"cycles_per_iter\": %.1f, \"ld_bytes_per_clk_per_sm\": %.1f, \"st_bytes_per_clk_per_sm\": %.1f, \"err\": \"%s\"}\n",
           name, warps, avg / iters, bytes_ld / (avg / iters), bytes_st / (avg / iters), cudaGetErrorString(e));
    cudaFree(out); cudaFree(cyc);
}

int main() {
    for (int w : {1, 4, 8}) {
        run<0>("tcgen05.ld 32x32b.x32 (4 per iter, wait each)", w);
        run<1>("tcgen05.st 32x32b.x32 (4 per iter)", w);
        run<2>("softmax pattern: 4 ld + 2 st per iter", w);
    }
    return 0;
}
