// SM clock as a function of time while OTHER kernels run: one thread of one tiny block (no shared memory, a handful of registers —
// it co-resides with the 1-CTA-per-SM GEMM / attention blocks) samples (%globaltimer, clock64) every `period_ns` and stores the
// pairs; clock between two samples = d(cycles) / d(ns).  Launched on its own stream before the work to observe
// (tools/clock_trace.py).  Measurement tool only (tools/bin/libattn_exp.so), not part of the product library.
#include <cuda_runtime.h>
#include <stdint.h>

__global__ void clock_probe_kernel(unsigned long long* __restrict__ buf, int n, unsigned period_ns, volatile int* stop) {
    for (int i = 0; i < n; ++i) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        const unsigned long long c = clock64();
        buf[2 * i] = t;
        buf[2 * i + 1] = c;
        if (*stop) {
            for (int j = i + 1; j < n; ++j) buf[2 * j] = 0;
            return;
        }
        __nanosleep(period_ns);
    }
}

extern "C" __attribute__((visibility("default"))) int s2v_clock_probe(void* buf_u64x2, int32_t n, uint32_t period_ns, void* stop_flag, void* stream) {
    // An SM keeps the L1 / shared-memory split of the blocks resident on it: with the default (small) carve-out of this kernel its SM
    // could not take a 200 KB GEMM / attention block while the probe runs (measured: persistent GEMMs ran on 147 SMs + a straggler
    // and took twice as long).  Ask for the maximum shared-memory carve-out so that the probe's SM stays usable.
    cudaFuncSetAttribute(clock_probe_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    clock_probe_kernel<<<1, 1, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<unsigned long long*>(buf_u64x2), n, period_ns,
                                                                       static_cast<volatile int*>(stop_flag));
    return (int)cudaPeekAtLastError();
}
