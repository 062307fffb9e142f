// T5 v1.1 encoder pieces around the tcgen05 GEMMs (the prompt encoder that runs once per video before the loop;
// transformers' modeling_t5.py T5LayerNorm / T5Attention / T5DenseGatedActDense — the reference imports that library,
// S/inference.py:13,185).  226 tokens per prompt: everything here is latency / weight-bandwidth bound, so these are plain coalesced
// kernels with warp-shuffle reductions; the projections go through s2v_linear.
#include <math.h>

#include <type_traits>

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

// The same function on warp-level tensor-core MMAs (mma.sync m16n8k16 bf16 -> fp32) for S <= 256: the scalar kernel above spends
// ~3000 instructions per query row (3.4 GFLOP on the FMA pipe = 0.128 ms per layer, 38 % of a T5-XXL encode).  One CTA = 64 query rows of one
// (batch, head), 4 warps x 16 rows; K and V of the head are staged once per CTA in shared memory with 72-half rows (ldmatrix reads 8 rows
// x 16 bytes conflict-free); a warp holds its 16 x NK score tile in registers (NK/8 accumulator tiles), applies the reference's rounding
// points (bf16(q k^T), bf16 add of the bias, fp32 softmax, P rounded to bf16) and feeds the accumulator registers straight back as the A
// fragments of P V (two adjacent 8-key tiles = one 16-key A fragment); V is read with ldmatrix.trans.  Single pass: no online softmax.
constexpr int T5M_LD = 72, T5M_THREADS = 128, T5M_ROWS = 64;
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const void* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}

template <int NT>   // NT = 8-key score tiles per row (even): S <= 8 * NT
__global__ void __launch_bounds__(T5M_THREADS) t5_attention_mma_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ bias, bf16* __restrict__ out,
                                                                       int S, int H) {
    extern __shared__ __align__(16) uint8_t t5_smem[];
    constexpr int NK = 8 * NT;
    bf16* sK = reinterpret_cast<bf16*>(t5_smem);
    bf16* sV = sK + NK * T5M_LD;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const int h = blockIdx.y, b = blockIdx.z;
    const long long ld = 3ll * H * 64;
    const bf16* base = qkv + (long long)b * S * ld + h * 64;
    for (int i = threadIdx.x; i < NK * 8; i += T5M_THREADS) {          // rows >= S are zero: their scores are masked, their P is 0
        const int j = i >> 3, c = i & 7;
        uint4 kv = make_uint4(0u, 0u, 0u, 0u), vv = kv;
        if (j < S) {
            kv = __ldg(reinterpret_cast<const uint4*>(base + (long long)j * ld + H * 64) + c);
            vv = __ldg(reinterpret_cast<const uint4*>(base + (long long)j * ld + 2 * H * 64) + c);
        }
        *reinterpret_cast<uint4*>(sK + j * T5M_LD + c * 8) = kv;
        *reinterpret_cast<uint4*>(sV + j * T5M_LD + c * 8) = vv;
    }
    // this thread's rows of the warp's 16-row tile: r0 = row g, r1 = row g + 8
    const int row0 = blockIdx.x * T5M_ROWS + warp * 16 + g, row1 = row0 + 8;
    uint32_t qa[4][4];          // A fragments of Q for the 4 k-steps of head_dim 64
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
        const int c = ks * 16 + 2 * t;
        qa[ks][0] = row0 < S ? __ldg(reinterpret_cast<const uint32_t*>(base + (long long)row0 * ld + c)) : 0u;
        qa[ks][1] = row1 < S ? __ldg(reinterpret_cast<const uint32_t*>(base + (long long)row1 * ld + c)) : 0u;
        qa[ks][2] = row0 < S ? __ldg(reinterpret_cast<const uint32_t*>(base + (long long)row0 * ld + c + 8)) : 0u;
        qa[ks][3] = row1 < S ? __ldg(reinterpret_cast<const uint32_t*>(base + (long long)row1 * ld + c + 8)) : 0u;
    }
    __syncthreads();
    if (blockIdx.x * T5M_ROWS + warp * 16 >= S) return;     // (after the barrier: the whole warp's tile lies beyond S)

    // ---- scores: S[16 x NK] = Q K^T
    float sc[NT][4];
#pragma unroll
    for (int n = 0; n < NT; ++n) {
        sc[n][0] = sc[n][1] = sc[n][2] = sc[n][3] = 0.f;
        // ldmatrix: lanes 0-7 address rows of matrix 0, 8-15 matrix 1, ...: matrices = 8-column chunks (k 0-7, 8-15, 16-23, 24-31 | 32-63)
        uint32_t kb0[4], kb1[4];
        const bf16* kp = sK + (n * 8 + (lane & 7)) * T5M_LD + (lane >> 3) * 8;
        ldmatrix_x4(kb0, kp);
        ldmatrix_x4(kb1, kp + 32);
        mma_bf16_16816(sc[n], qa[0], kb0[0], kb0[1]);
        mma_bf16_16816(sc[n], qa[1], kb0[2], kb0[3]);
        mma_bf16_16816(sc[n], qa[2], kb1[0], kb1[1]);
        mma_bf16_16816(sc[n], qa[3], kb1[2], kb1[3]);
    }
    // ---- bf16(q k^T) + bias in bf16, row max.  c0,c1 = (row0, cols 8n + 2t, +1), c2,c3 = (row1, same cols)
    const bf16* b0p = bias + ((long long)h * S + (row0 < S ? row0 : 0)) * S;
    const bf16* b1p = bias + ((long long)h * S + (row1 < S ? row1 : 0)) * S;
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int n = 0; n < NT; ++n) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int j = n * 8 + 2 * t + e;
            if (j < S) {
                sc[n][e] = bf16r(bf16r(sc[n][e]) + __bfloat162float(b0p[j]));
                sc[n][2 + e] = bf16r(bf16r(sc[n][2 + e]) + __bfloat162float(b1p[j]));
            } else {
                sc[n][e] = -INFINITY;
                sc[n][2 + e] = -INFINITY;
            }
            mx0 = fmaxf(mx0, sc[n][e]);
            mx1 = fmaxf(mx1, sc[n][2 + e]);
        }
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
    for (int n = 0; n < NT; ++n) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            sc[n][e] = __expf(sc[n][e] - mx0);            // exp(-inf) = 0 for the masked columns
            sc[n][2 + e] = __expf(sc[n][2 + e] - mx1);
            sum0 += sc[n][e];
            sum1 += sc[n][2 + e];
        }
    }
    sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1); sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
    sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1); sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);
    const float inv0 = 1.0f / sum0, inv1 = 1.0f / sum1;

    // ---- O[16 x 64] = P V: A fragment of key block kk = the accumulators of score tiles 2kk and 2kk + 1, rounded to bf16
    float oc[8][4];
#pragma unroll
    for (int n = 0; n < 8; ++n) oc[n][0] = oc[n][1] = oc[n][2] = oc[n][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < NT / 2; ++kk) {
        uint32_t pa[4];
        pa[0] = pack_bf16x2(sc[2 * kk][0] * inv0, sc[2 * kk][1] * inv0);
        pa[1] = pack_bf16x2(sc[2 * kk][2] * inv1, sc[2 * kk][3] * inv1);
        pa[2] = pack_bf16x2(sc[2 * kk + 1][0] * inv0, sc[2 * kk + 1][1] * inv0);
        pa[3] = pack_bf16x2(sc[2 * kk + 1][2] * inv1, sc[2 * kk + 1][3] * inv1);
        // V[16 keys x 64 channels] as B fragments: ldmatrix.trans, matrices = (keys 0-7 | 8-15) x (channels 16m .. 16m + 7 | + 8 .. + 15)
        const bf16* vp = sV + (kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * T5M_LD + (lane >> 4) * 8;
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            uint32_t vb[4];
            ldmatrix_x4_trans(vb, vp + m * 16);
            mma_bf16_16816(oc[2 * m], pa, vb[0], vb[1]);
            mma_bf16_16816(oc[2 * m + 1], pa, vb[2], vb[3]);
        }
    }
    bf16* o0 = out + ((long long)b * S + row0) * (H * 64) + h * 64;
    bf16* o1 = out + ((long long)b * S + row1) * (H * 64) + h * 64;
#pragma unroll
    for (int n = 0; n < 8; ++n) {
        if (row0 < S) *reinterpret_cast<uint32_t*>(o0 + n * 8 + 2 * t) = pack_bf16x2(oc[n][0], oc[n][1]);
        if (row1 < S) *reinterpret_cast<uint32_t*>(o1 + n * 8 + 2 * t) = pack_bf16x2(oc[n][2], oc[n][3]);
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
    if (S <= 256) {      // tensor-core form (single pass, the score tile of a 16-row warp tile in registers)
        dim3 gridm((S + T5M_ROWS - 1) / T5M_ROWS, H, B);
        auto launch = [&](auto nt_tag) -> int {
            constexpr int NT = decltype(nt_tag)::value;
            const int smem_m = 2 * 8 * NT * T5M_LD * 2;
            auto kern = t5_attention_mma_kernel<NT>;
            int r = ensure_smem_optin(reinterpret_cast<const void*>(kern), smem_m, "cudaFuncSetAttribute(t5_attention_mma)");
            if (r) return r;
            kern<<<gridm, T5M_THREADS, smem_m, stream>>>(static_cast<const bf16*>(qkv), static_cast<const bf16*>(bias), static_cast<bf16*>(out), S, H);
            return check_launch("t5_attention_mma_kernel");
        };
        if (S <= 64) return launch(std::integral_constant<int, 8>{});
        if (S <= 128) return launch(std::integral_constant<int, 16>{});
        if (S <= 240) return launch(std::integral_constant<int, 30>{});
        return launch(std::integral_constant<int, 32>{});
    }
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
