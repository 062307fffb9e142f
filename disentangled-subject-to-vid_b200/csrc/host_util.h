// Host-side helpers shared by the C-ABI translation units: error reporting, device checks, TMA descriptor encoding.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace s2v {

int set_error(int code, const char* msg);                   // records msg (thread local), returns code
int set_cuda_error(cudaError_t e, const char* where);       // returns (int)e (> 0)
int check_launch(const char* kernel_name);                  // cudaPeekAtLastError after a launch
int ensure_device();                                        // 0 if the current device is sm_100, else S2V_E_NO_DEVICE
int sm_count();                                             // multiprocessors of the current device (cached)
// Opt a kernel in to `bytes` of dynamic shared memory on the CURRENT device.  The attribute is per device (context), so the
// "done" flag is kept per (kernel, device ordinal) — a process that drives several GPUs opts in on each of them.
int ensure_smem_optin(const void* kernel, int bytes, const char* what);

// 2D row-major bf16 tensor [rows, cols] with leading dimension `ld` (elements); box = {box_cols, box_rows},
// 128-byte swizzle, zero fill out of bounds.
int make_tmap_2d_bf16(CUtensorMap* out, const void* base, int64_t cols, int64_t rows, int64_t ld, int box_cols,
                      int box_rows);
// generic up-to-4D bf16 map: dims/strides innermost first (strides in BYTES for dims 1..rank-1)
int make_tmap_nd_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                      const uint32_t* box, int swizzle_bytes = 128);

}  // namespace s2v
