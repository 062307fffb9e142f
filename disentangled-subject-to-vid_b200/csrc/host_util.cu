#include "host_util.h"

#include <stdio.h>
#include <string.h>

#include <mutex>

#include "s2v_b200.h"

namespace s2v {

static thread_local char g_err[512] = "";

int set_error(int code, const char* msg) {
    snprintf(g_err, sizeof(g_err), "%s", msg);
    return code;
}

int set_cuda_error(cudaError_t e, const char* where) {
    snprintf(g_err, sizeof(g_err), "%s: %s (%s)", where, cudaGetErrorName(e), cudaGetErrorString(e));
    return static_cast<int>(e);
}

int check_launch(const char* kernel_name) {
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) {
        cudaGetLastError();
        return set_cuda_error(e, kernel_name);
    }
    return 0;
}

static int g_sm_count[64];
static int g_dev_ok[64];  // 0 unknown, 1 ok, -1 not sm_100

int ensure_device() {
    int dev = -1;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess || dev < 0) {
        cudaGetLastError();
        return set_error(S2V_E_NO_DEVICE, "no CUDA device available (s2v_b200 has no CPU fallback)");
    }
    if (dev >= 64) return set_error(S2V_E_NO_DEVICE, "device ordinal out of range");
    if (g_dev_ok[dev] == 0) {
        cudaDeviceProp prop;
        e = cudaGetDeviceProperties(&prop, dev);
        if (e != cudaSuccess) return set_cuda_error(e, "cudaGetDeviceProperties");
        g_sm_count[dev] = prop.multiProcessorCount;
        g_dev_ok[dev] = (prop.major == 10) ? 1 : -1;
    }
    if (g_dev_ok[dev] < 0)
        return set_error(S2V_E_NO_DEVICE, "device is not sm_100 (B200); s2v_b200 ships sm_100a kernels only");
    return 0;
}

int sm_count() {
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || g_sm_count[dev] == 0) return 148;
    return g_sm_count[dev];
}

int ensure_smem_optin(const void* kernel, int bytes, const char* what) {
    // per (kernel, device): the largest dynamic shared-memory size opted into so far (a kernel whose footprint depends on the call —
    // the scalar T5 attention: 2 x S rows — asks again when it needs more)
    struct Entry { const void* k; int dev; int bytes; };
    static Entry table[256];
    static int n = 0;
    static std::mutex mu;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return set_error(S2V_E_NO_DEVICE, "no current CUDA device");
    std::lock_guard<std::mutex> lock(mu);
    Entry* e = nullptr;
    for (int i = 0; i < n; ++i)
        if (table[i].k == kernel && table[i].dev == dev) e = &table[i];
    if (e && e->bytes >= bytes) return 0;
    if (!e) {
        if (n == 256) return set_error(S2V_E_DRIVER, "ensure_smem_optin: kernel table full");
        e = &table[n++];
        e->k = kernel;
        e->dev = dev;
        e->bytes = 0;
    }
    cudaError_t err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (err != cudaSuccess) return set_cuda_error(err, what);
    e->bytes = bytes;
    return 0;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
        if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) {
            cudaGetLastError();
            return nullptr;
        }
        fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

int make_tmap_nd_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                      const uint32_t* box, int swizzle_bytes) {
    EncodeTiledFn fn = get_encode();
    if (!fn) return set_error(S2V_E_DRIVER, "cuTensorMapEncodeTiled entry point not available");
    if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) return set_error(S2V_E_BADARG, "TMA base must be 16-byte aligned");
    cuuint64_t gdim[5];
    cuuint64_t gstr[5];
    cuuint32_t bdim[5];
    cuuint32_t estr[5];
    for (int i = 0; i < rank; ++i) {
        gdim[i] = dims[i];
        bdim[i] = box[i];
        estr[i] = 1;
        if (i > 0) {
            if (strides_bytes[i] % 16) return set_error(S2V_E_BADARG, "TMA strides must be multiples of 16 bytes");
            gstr[i - 1] = strides_bytes[i];
        }
    }
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(base), gdim, gstr, bdim, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        char buf[128];
        snprintf(buf, sizeof(buf), "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
        return set_error(S2V_E_DRIVER, buf);
    }
    return 0;
}

int make_tmap_2d_bf16(CUtensorMap* out, const void* base, int64_t cols, int64_t rows, int64_t ld, int box_cols,
                      int box_rows) {
    const uint64_t dims[2] = {(uint64_t)cols, (uint64_t)rows};
    const uint64_t strides[2] = {2, (uint64_t)ld * 2};
    const uint32_t box[2] = {(uint32_t)box_cols, (uint32_t)box_rows};
    return make_tmap_nd_bf16(out, base, 2, dims, strides, box);
}

}  // namespace s2v

extern "C" int s2v_abi_version(void) { return S2V_ABI_VERSION; }
extern "C" const char* s2v_last_error(void) { return s2v::g_err; }
extern "C" int s2v_device_check(int dev) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || dev < 0 || dev >= n) {
        cudaGetLastError();
        return s2v::set_error(S2V_E_NO_DEVICE, "no such CUDA device");
    }
    cudaDeviceProp prop;
    cudaError_t e = cudaGetDeviceProperties(&prop, dev);
    if (e != cudaSuccess) return s2v::set_cuda_error(e, "cudaGetDeviceProperties");
    if (prop.major != 10) return s2v::set_error(S2V_E_NO_DEVICE, "device is not sm_100 (B200)");
    return 0;
}

extern "C" int64_t s2v_workspace_bytes(int32_t B, int32_t S, int32_t D, int32_t ff_dim, int32_t lora_cols, int32_t n_mod) {
    if (B <= 0 || S <= 0 || D <= 0 || ff_dim <= 0 || lora_cols < 0 || n_mod <= 0) return s2v::set_error(S2V_E_BADARG, "s2v_workspace_bytes: bad geometry");
    auto up = [](int64_t b) { return (b + 255) & ~int64_t(255); };
    const int64_t rows = int64_t(B) * S;
    const int64_t lt_cols = lora_cols > 8 ? lora_cols : 8;
    return 3 * up(rows * D * 2) + up(rows * 3 * D * 2) + up(rows * ff_dim * 2) + up(rows * lt_cols * 2) + up(int64_t(n_mod) * B * 6 * D * 4);
}
