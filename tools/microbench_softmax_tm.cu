// Softmax inner loop WITH its TMEM traffic, as a stand-alone proxy of the attention kernel's softmax warps (the attention
// kernel is entirely softmax-side bound: it takes 10.2 ms with the MMAs removed vs 10.6 ms with them).
// Each thread owns E scores of a tile (E = 64: one thread per row of a 64-key tile; E = 32: two threads per row).
//   PIPE 0: ld, wait, compute, st, wait            (the shipped kernel)
//   PIPE 1: ld(t+1) issued before compute(t) into a second register set (unrolled by two)
//   PIPE 2: PIPE 0 but the tcgen05.wait::st is deferred to just before the next store
//   PIPE 3: PIPE 1 + deferred wait::st
// Output: SM cycles per 64 keys x 256 rows (the tensor core needs 512; the exponential unit 896 at POLY16 = 1).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/mb_softmax_tm tools/microbench_softmax_tm.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t pack2(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void unpack2(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) { uint64_t d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float fmax3(float a, float b, float c) { float d; asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ uint32_t cvt2(float lo, float hi) { uint32_t r; asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo)); return r; }
__device__ __forceinline__ void exp2_poly2(uint64_t X, float& r0, float& r1, float& xmax) {
    const float MAGIC = 12582912.0f;
    float x0, x1;
    unpack2(X, x0, x1);
    xmax = fmax3(xmax, x0, x1);
    x0 = fmaxf(x0, -125.0f); x1 = fmaxf(x1, -125.0f);
    X = pack2(x0, x1);
    const uint64_t T = fadd2(X, pack2(MAGIC, MAGIC));
    const uint64_t NF = fadd2(T, pack2(-MAGIC, -MAGIC));
    const uint64_t F = ffma2(NF, pack2(-1.0f, -1.0f), X);
    uint64_t P = ffma2(pack2(0.05517166f, 0.05517166f), F, pack2(0.24261112f, 0.24261112f));
    P = ffma2(P, F, pack2(0.69326099f, 0.69326099f));
    P = ffma2(P, F, pack2(0.99992807f, 0.99992807f));
    float p0, p1, t0, t1;
    unpack2(P, p0, p1); unpack2(T, t0, t1);
    r0 = __uint_as_float(__float_as_uint(p0) + (__float_as_uint(t0) << 23));
    r1 = __uint_as_float(__float_as_uint(p1) + (__float_as_uint(t1) << 23));
}
#define LD32(taddr, r) asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];" \
    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), \
      "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]) : "r"(taddr) : "memory")
#define ST16(taddr, r) asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" \
    :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory")
#define WAIT_LD() asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory")
#define WAIT_ST() asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory")

// probabilities of 32 scores -> pk[16]; POLY16 pairs of every 8 pairs on the FMA pipe
template <int POLY16>
__device__ __forceinline__ void half_tile(const uint32_t (&sv)[32], uint32_t (&pk)[16], uint64_t C2, uint64_t M2, uint64_t& acc0, uint64_t& acc1, float& xmax) {
#pragma unroll
    for (int i = 0; i < 32; i += 8) {
#pragma unroll
        for (int h = 0; h < 4; ++h) {
            const uint64_t X = ffma2(pack2(__uint_as_float(sv[i + 2 * h]), __uint_as_float(sv[i + 2 * h + 1])), C2, M2);
            float p0, p1;
            if (((i >> 3) & 1) * 4 + h < POLY16) {
                exp2_poly2(X, p0, p1, xmax);
            } else {
                float x0, x1;
                unpack2(X, x0, x1);
                p0 = ex2(x0); p1 = ex2(x1);
            }
            if (h & 1) acc1 = fadd2(acc1, pack2(p0, p1)); else acc0 = fadd2(acc0, pack2(p0, p1));
            pk[i / 2 + h] = cvt2(p0, p1);
        }
    }
}

template <int WARPS, int E, int PIPE, int POLY16>
__global__ void __launch_bounds__(WARPS * 32, 1) bench(float* out, long long* cycles, int iters, float c, float m0) {
    __shared__ uint32_t tm_slot;
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&tm_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int w = threadIdx.x >> 5;
    const uint32_t tS = tm_slot + (uint32_t((w & 3) * 32) << 16) + (w >> 2) * (256 / (WARPS / 4));   // score columns of this warp
    const uint32_t tP = tm_slot + (uint32_t((w & 3) * 32) << 16) + 256 + (w >> 2) * 32;
    {   // fill the score region with finite values
        uint32_t z[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) z[i] = __float_as_uint(-2.0f + 0.01f * ((threadIdx.x + i) & 63));
        for (int k = 0; k < 128 / 16 && k * 16 < 256 / (WARPS / 4); ++k) ST16(tS + k * 16, z);
        WAIT_ST();
    }
    float lsum = 0.f, mneg = m0, xmax = -1e30f;
    const uint64_t C2 = pack2(c, c);
    uint32_t a0[32], a1[32], b0[32], b1[32];
    uint32_t pk[16];
    auto load = [&](uint32_t (&lo)[32], uint32_t (&hi)[32], int t) {
        LD32(tS + (t & 1) * (E == 64 ? 64 : 32), lo);
        if (E == 64) LD32(tS + (t & 1) * 64 + 32, hi);
    };
    auto compute_store = [&](uint32_t (&lo)[32], uint32_t (&hi)[32]) {
        const uint64_t M2 = pack2(mneg, mneg);
        uint64_t acc0 = 0, acc1 = 0;
        half_tile<POLY16>(lo, pk, C2, M2, acc0, acc1, xmax);
        if (PIPE >= 2) WAIT_ST();
        ST16(tP, pk);
        if (E == 64) {
            half_tile<POLY16>(hi, pk, C2, M2, acc0, acc1, xmax);
            ST16(tP + 16, pk);
        }
        if (PIPE < 2) WAIT_ST();
        float x, y, z, u;
        unpack2(acc0, x, y); unpack2(acc1, z, u);
        lsum += (x + y) + (z + u);
        mneg += 1e-6f;
    };
    __syncthreads();
    const long long t0 = clock64();
    if (PIPE == 0 || PIPE == 2) {
#pragma unroll 1
        for (int it = 0; it < iters; ++it) {
            load(a0, a1, it);
            WAIT_LD();
            compute_store(a0, a1);
        }
    } else {
        load(a0, a1, 0);
#pragma unroll 1
        for (int it = 0; it < iters; it += 2) {
            WAIT_LD();
            load(b0, b1, it + 1);
            compute_store(a0, a1);
            WAIT_LD();
            load(a0, a1, it + 2);
            compute_store(b0, b1);
        }
        WAIT_LD();
    }
    WAIT_ST();
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = lsum + mneg + xmax + __uint_as_float(a0[3] ^ b0[5]);
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm_slot), "r"(512) : "memory");
}

template <int WARPS, int E, int PIPE, int POLY16>
void run() {
    const int blocks = 148, iters = 4000;
    float* out;
    long long* cyc;
    cudaMalloc(&out, blocks * WARPS * 32 * sizeof(float));
    cudaMalloc(&cyc, blocks * sizeof(long long));
    for (int r = 0; r < 2; ++r) bench<WARPS, E, PIPE, POLY16><<<blocks, WARPS * 32>>>(out, cyc, iters, 0.18f, -0.5f);
    cudaError_t e = cudaDeviceSynchronize();
    long long hc[148];
    cudaMemcpy(hc, cyc, sizeof(hc), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < blocks; ++i) avg += hc[i];
    avg /= blocks;
    // one iteration of all warps = WARPS*32*E scores; a 64-key step of a 256-row CTA = 16384 scores
    const double per_step = avg / iters * 16384.0 / (WARPS * 32.0 * E);
    printf("{\"warps\": %d, \"scores_per_thread\": %d, \"pipe\": %d, \"poly16\": %d, \"cycles_per_64key_step\": %.1f, \"err\": \"%s\"}\n", WARPS, E, PIPE, POLY16,
           per_step, cudaGetErrorString(e));
    cudaFree(out); cudaFree(cyc);
}

int main() {
    run<8, 64, 0, 1>(); run<8, 64, 1, 1>(); run<8, 64, 2, 1>(); run<8, 64, 3, 1>();
    run<8, 64, 0, 2>(); run<8, 64, 1, 2>(); run<8, 64, 3, 2>(); run<8, 64, 0, 0>(); run<8, 64, 3, 0>();
    run<16, 32, 0, 1>(); run<16, 32, 1, 1>(); run<16, 32, 3, 1>(); run<16, 32, 3, 2>();
    return 0;
}
