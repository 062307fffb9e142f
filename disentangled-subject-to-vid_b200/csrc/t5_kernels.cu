// T5 v1.1 encoder pieces around the tcgen05 GEMMs (the prompt encoder that runs once per video before the loop;
// transformers' modeling_t5.py T5LayerNorm / T5Attention / T5DenseGatedActDense — the reference imports that library,
// S/inference.py:13,185).  226 tokens per prompt: everything here is latency / weight-bandwidth bound, so these are plain coalesced
// kernels with warp-shuffle reductions; the projections go through s2v_linear.
#include <math.h>

#include "common.cuh"
#include "host_util.h"
#include "s2v_b200.h"

namespace s2v {

__device__ __forceinline__ float t5_warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float t5_warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float bf16r(float x) { return __bfloat162float(__float2bfloat16(x)); }

// out[i, :] = table[ids[i], :]   (nn.Embedding: shared.weight)
__global__ void __launch_bounds__(256) gather_rows_kernel(const bf16* __restrict__ table, const long long* __restrict__ ids, bf16* __restrict__ out,
                                                          int n, int D, int V) {
    const int vec = D / 8;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < (long long)n * vec; i += (long long)gridDim.x * blockDim.x) {
        const int r = int(i / vec), c = int(i - (long long)r * vec);
        long long id = ids[r];
        id = id < 0 ? 0 : (id >= V ? V - 1 : id);
        reinterpret_cast<uint4*>(out)[i] = __ldg(reinterpret_cast<const uint4*>(table + id * D) + c);
    }
}

// T5LayerNorm: out = w * bf16(x * rsqrt(mean(x^2) + eps)), the variance in fp32, both roundings as modeling_t5.py does them.
// One warp per row; D <= 8192 (the row is re-read from L1/L2 for the second pass, it is 8 KB).
__global__ void __launch_bounds__(256) rmsnorm_kernel(const bf16* __restrict__ x, const bf16* __restrict__ w, bf16* __restrict__ out, int rows, int D,
                                                      float eps) {
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= rows) return;
    const uint4* xr = reinterpret_cast<const uint4*>(x + (long long)row * D);
    const int nvec = D / 8;
    float ss = 0.f;
    for (int v = lane; v < nvec; v += 32) {
        const uint4 u = __ldg(xr + v);
        const uint32_t ww[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float a = bf16_lo(ww[j]), b = bf16_hi(ww[j]);
            ss = fmaf(a, a, ss);
            ss = fmaf(b, b, ss);
        }
    }
    const float rstd = rsqrtf(t5_warp_sum(ss) / float(D) + eps);
    uint4* orow = reinterpret_cast<uint4*>(out + (long long)row * D);
    for (int v = lane; v < nvec; v += 32) {
        const uint4 u = __ldg(xr + v), g = __ldg(reinterpret_cast<const uint4*>(w) + v);
        const uint32_t xw[4] = {u.x, u.y, u.z, u.w}, gw[4] = {g.x, g.y, g.z, g.w};
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j)
            o[j] = pack_bf16x2(bf16_lo(gw[j]) * bf16r(bf16_lo(xw[j]) * rstd), bf16_hi(gw[j]) * bf16r(bf16_hi(xw[j]) * rstd));
        orow[v] = make_uint4(o[0], o[1], o[2], o[3]);
    }
}

// T5DenseGatedActDense: out[m, f] = bf16(gelu_new(g[m, f])) * g[m, F + f]  (wi_0 | wi_1 stacked along N in one GEMM)
__global__ void __launch_bounds__(256) gated_gelu_kernel(const bf16* __restrict__ g, bf16* __restrict__ out, long long M, int F) {
    const int vec = F / 8;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < M * vec; i += (long long)gridDim.x * blockDim.x) {
        const long long m = i / vec;
        const int c = int(i - m * vec);
        const uint4 a = __ldg(reinterpret_cast<const uint4*>(g + m * 2 * F) + c);
        const uint4 b = __ldg(reinterpret_cast<const uint4*>(g + m * 2 * F + F) + c);
        const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) o[j] = pack_bf16x2(bf16r(gelu_tanh(bf16_lo(aw[j]))) * bf16_lo(bw[j]), bf16r(gelu_tanh(bf16_hi(aw[j]))) * bf16_hi(bw[j]));
        reinterpret_cast<uint4*>(out)[i] = make_uint4(o[0], o[1], o[2], o[3]);
    }
}

// T5Attention (encoder self-attention, head_dim 64, NO 1/sqrt(d) scaling, additive relative-position bias, no mask):
//   scores = bf16(q k^T) + bias  (bf16 add, like the bf16 tensors of modeling_t5.py), softmax in fp32, P rounded to bf16, out = P v.
// One CTA = 16 query rows of one (batch, head): K and V of the head are staged in shared memory once (rows padded to 66 halves:
// lane j reads key row j at bank (j + d/2) % 32 — conflict free), a warp owns a query row at a time: lane <-> key for the scores,
// lane <-> output dims (2l, 2l+1) for P v.
constexpr int T5A_ROWS = 16, T5A_THREADS = 256, T5A_LD = 66;
template <int KPL>   // keys per lane: S <= 32 * KPL
__global__ void __launch_bounds__(T5A_THREADS) t5_attention_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ bias, bf16* __restrict__ out,
                                                                   int S, int H) {
    extern __shared__ uint8_t t5_smem[];
    bf16* sK = reinterpret_cast<bf16*>(t5_smem);
    bf16* sV = sK + (size_t)S * T5A_LD;
    float* sP = reinterpret_cast<float*>(sV + (size_t)S * T5A_LD);     // [8 warps][32 * KPL]
    float* sQ = sP + 8 * 32 * KPL;                                     // [8 warps][64]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int h = blockIdx.y, b = blockIdx.z;
    const long long ld = 3ll * H * 64;
    const bf16* base = qkv + (long long)b * S * ld + h * 64;
    for (int i = threadIdx.x; i < S * 8; i += T5A_THREADS) {           // 8 x 16-byte vectors per 64-wide row
        const int j = i >> 3, c = i & 7;
        const uint4 kv = __ldg(reinterpret_cast<const uint4*>(base + (long long)j * ld + H * 64) + c);
        const uint4 vv = __ldg(reinterpret_cast<const uint4*>(base + (long long)j * ld + 2 * H * 64) + c);
        uint32_t* dk = reinterpret_cast<uint32_t*>(sK + (size_t)j * T5A_LD + c * 8);
        uint32_t* dv = reinterpret_cast<uint32_t*>(sV + (size_t)j * T5A_LD + c * 8);
        dk[0] = kv.x; dk[1] = kv.y; dk[2] = kv.z; dk[3] = kv.w;
        dv[0] = vv.x; dv[1] = vv.y; dv[2] = vv.z; dv[3] = vv.w;
    }
    __syncthreads();
    float* myP = sP + warp * 32 * KPL;
    float* myQ = sQ + warp * 64;
    for (int r = warp; r < T5A_ROWS; r += 8) {
        const int i = blockIdx.x * T5A_ROWS + r;
        if (i >= S) break;
        const uint32_t qw = __ldg(reinterpret_cast<const uint32_t*>(base + (long long)i * ld) + lane);
        myQ[2 * lane] = bf16_lo(qw);
        myQ[2 * lane + 1] = bf16_hi(qw);
        __syncwarp();
        float s[KPL];
        float mx = -INFINITY;
        const bf16* brow = bias + ((long long)h * S + i) * S;
#pragma unroll
        for (int t = 0; t < KPL; ++t) {
            const int j = lane + 32 * t;
            s[t] = -INFINITY;
            if (j < S) {
                const uint32_t* kr = reinterpret_cast<const uint32_t*>(sK + (size_t)j * T5A_LD);
                float acc = 0.f;
#pragma unroll
                for (int d2 = 0; d2 < 32; ++d2) {
                    const uint32_t kk = kr[d2];
                    acc = fmaf(myQ[2 * d2], bf16_lo(kk), acc);
                    acc = fmaf(myQ[2 * d2 + 1], bf16_hi(kk), acc);
                }
                s[t] = bf16r(bf16r(acc) + __bfloat162float(brow[j]));
                mx = fmaxf(mx, s[t]);
            }
        }
        mx = t5_warp_max(mx);
        float sum = 0.f;
#pragma unroll
        for (int t = 0; t < KPL; ++t) {
            s[t] = (lane + 32 * t < S) ? __expf(s[t] - mx) : 0.f;
            sum += s[t];
        }
        const float inv = 1.0f / t5_warp_sum(sum);
#pragma unroll
        for (int t = 0; t < KPL; ++t) myP[lane + 32 * t] = bf16r(s[t] * inv);
        __syncwarp();
        float o0 = 0.f, o1 = 0.f;
        for (int j = 0; j < S; ++j) {
            const float pj = myP[j];
            const uint32_t vv = reinterpret_cast<const uint32_t*>(sV + (size_t)j * T5A_LD)[lane];
            o0 = fmaf(pj, bf16_lo(vv), o0);
            o1 = fmaf(pj, bf16_hi(vv), o1);
        }
        reinterpret_cast<uint32_t*>(out + ((long long)b * S + i) * (H * 64) + h * 64)[lane] = pack_bf16x2(o0, o1);
        __syncwarp();
    }
}

static inline int t5_grid(long long n) {
    long long b = (n + 255) / 256;
    const long long cap = (long long)sm_count() * 16;
    return (int)(b > cap ? cap : (b < 1 ? 1 : b));
}

}  // namespace s2v

using namespace s2v;

extern "C" int s2v_gather_rows(const void* table, const int64_t* ids, void* out, int32_t n, int32_t D, int32_t V, void* stream) {
    if (!table || !ids || !out) return set_error(S2V_E_BADARG, "s2v_gather_rows: null pointer");
    if (n <= 0 || D <= 0 || V <= 0 || (D % 8)) return set_error(S2V_E_BADARG, "s2v_gather_rows: bad shape (D % 8 != 0?)");
    int rc = ensure_device();
    if (rc) return rc;
    gather_rows_kernel<<<t5_grid((long long)n * (D / 8)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const bf16*>(table), reinterpret_cast<const long long*>(ids), static_cast<bf16*>(out), n, D, V);
    return check_launch("gather_rows_kernel");
}

extern "C" int s2v_rmsnorm(const void* x, const void* w, void* out, int32_t rows, int32_t D, float eps, void* stream) {
    if (!x || !w || !out) return set_error(S2V_E_BADARG, "s2v_rmsnorm: null pointer");
    if (rows <= 0 || D <= 0 || (D % 8)) return set_error(S2V_E_BADARG, "s2v_rmsnorm: bad shape (D % 8 != 0?)");
    int rc = ensure_device();
    if (rc) return rc;
    rmsnorm_kernel<<<(rows + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const bf16*>(x), static_cast<const bf16*>(w),
                                                                                 static_cast<bf16*>(out), rows, D, eps);
    return check_launch("rmsnorm_kernel");
}

extern "C" int s2v_gated_gelu(const void* g, void* out, int64_t M, int32_t F, void* stream) {
    if (!g || !out) return set_error(S2V_E_BADARG, "s2v_gated_gelu: null pointer");
    if (M <= 0 || F <= 0 || (F % 8)) return set_error(S2V_E_BADARG, "s2v_gated_gelu: bad shape (F % 8 != 0?)");
    int rc = ensure_device();
    if (rc) return rc;
    gated_gelu_kernel<<<t5_grid(M * (F / 8)), 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const bf16*>(g), static_cast<bf16*>(out), M, F);
    return check_launch("gated_gelu_kernel");
}

extern "C" int s2v_t5_attention(const void* qkv, const void* bias, void* out, int32_t B, int32_t S, int32_t H, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (!qkv || !bias || !out) return set_error(S2V_E_BADARG, "s2v_t5_attention: null pointer");
    if (B <= 0 || S <= 0 || H <= 0 || B > 65535 || H > 65535) return set_error(S2V_E_BADARG, "s2v_t5_attention: bad shape");
    if (S > 512) return set_error(S2V_E_UNSUPPORTED, "s2v_t5_attention: at most 512 tokens (the prompt encoder runs 226)");
    int rc = ensure_device();
    if (rc) return rc;
    const int kpl = S <= 256 ? 8 : 16;
    const int smem = 2 * S * T5A_LD * 2 + 8 * 32 * kpl * 4 + 8 * 64 * 4;
    dim3 grid((S + T5A_ROWS - 1) / T5A_ROWS, H, B);
    if (kpl == 8) {
        if ((rc = ensure_smem_optin(reinterpret_cast<const void*>(t5_attention_kernel<8>), smem, "cudaFuncSetAttribute(t5_attention)"))) return rc;
        t5_attention_kernel<8><<<grid, T5A_THREADS, smem, stream>>>(static_cast<const bf16*>(qkv), static_cast<const bf16*>(bias), static_cast<bf16*>(out), S, H);
    } else {
        if ((rc = ensure_smem_optin(reinterpret_cast<const void*>(t5_attention_kernel<16>), smem, "cudaFuncSetAttribute(t5_attention)"))) return rc;
        t5_attention_kernel<16><<<grid, T5A_THREADS, smem, stream>>>(static_cast<const bf16*>(qkv), static_cast<const bf16*>(bias), static_cast<bf16*>(out), S, H);
    }
    return check_launch("t5_attention_kernel");
}
