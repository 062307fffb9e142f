// HBM-bound kernels of the denoising step: AdaLN-Zero LayerNorm+modulate, final norms, q/k LayerNorm + RoPE,
// small-batch linears (timestep MLP / modulation vectors), patchify / unpatchify, CFG + DDIM update.
// All are coalesced 16-byte vector kernels with warp-shuffle reductions; statistics and arithmetic in fp32.
#include <math.h>
#include <stdlib.h>

#include <type_traits>

#include "common.cuh"
#include "host_util.h"
#include "s2v_b200.h"

namespace s2v {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
    f[0] = bf16_lo(u.x); f[1] = bf16_hi(u.x); f[2] = bf16_lo(u.y); f[3] = bf16_hi(u.y);
    f[4] = bf16_lo(u.z); f[5] = bf16_hi(u.z); f[6] = bf16_lo(u.w); f[7] = bf16_hi(u.w);
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
    uint4 u;
    u.x = pack_bf16x2(f[0], f[1]); u.y = pack_bf16x2(f[2], f[3]);
    u.z = pack_bf16x2(f[4], f[5]); u.w = pack_bf16x2(f[6], f[7]);
    return u;
}

// ------------------------------------------------------------------------------------------ LayerNorm (+ modulate)
// One warp per row; the row (D <= 32*8*MAXV elements) is held in registers so HBM sees exactly one read + one write.
// Two-pass statistics (mean, then centred sum of squares) in fp32, as ATen's LayerNorm.
template <int MAXV>
struct RowRegs {
    float v[MAXV][8];
};

template <int MAXV>
__device__ __forceinline__ void row_load(const bf16* __restrict__ x, int nvec, int lane, RowRegs<MAXV>& r) {
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
        const int vi = i * 32 + lane;
        if (vi < nvec) {
            const uint4 u = __ldg(reinterpret_cast<const uint4*>(x) + vi);
            unpack8(u, r.v[i]);
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) r.v[i][j] = 0.f;
        }
    }
}

template <int MAXV>
__device__ __forceinline__ void row_stats(const RowRegs<MAXV>& r, int nvec, int lane, int D, float eps, float& mean, float& rstd) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) s += r.v[i][j];
    mean = warp_sum(s) / float(D);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
        if (i * 32 + lane < nvec) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float d = r.v[i][j] - mean;
                q += d * d;
            }
        }
    }
    rstd = rsqrtf(warp_sum(q) / float(D) + eps);
}

// y = ((x - mean) * rstd * w + b)
template <int MAXV>
__device__ __forceinline__ void row_affine(RowRegs<MAXV>& r, int nvec, int lane, float mean, float rstd,
                                           const bf16* __restrict__ w, const bf16* __restrict__ b) {
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
        const int vi = i * 32 + lane;
        if (vi < nvec) {
            float wf[8], bfv[8];
            unpack8(__ldg(reinterpret_cast<const uint4*>(w) + vi), wf);
            unpack8(__ldg(reinterpret_cast<const uint4*>(b) + vi), bfv);
#pragma unroll
            for (int j = 0; j < 8; ++j) r.v[i][j] = (r.v[i][j] - mean) * rstd * wf[j] + bfv[j];
        }
    }
}

template <int MAXV>
__device__ __forceinline__ void row_modulate_store(const RowRegs<MAXV>& r, int nvec, int lane, const float* __restrict__ shift,
                                                   const float* __restrict__ scale, bf16* __restrict__ out) {
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
        const int vi = i * 32 + lane;
        if (vi < nvec) {
            const float4 s0 = __ldg(reinterpret_cast<const float4*>(scale) + 2 * vi);
            const float4 s1 = __ldg(reinterpret_cast<const float4*>(scale) + 2 * vi + 1);
            const float4 h0 = __ldg(reinterpret_cast<const float4*>(shift) + 2 * vi);
            const float4 h1 = __ldg(reinterpret_cast<const float4*>(shift) + 2 * vi + 1);
            const float sc[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
            const float sh[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
            float o[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = r.v[i][j] * (1.0f + sc[j]) + sh[j];
            reinterpret_cast<uint4*>(out)[vi] = pack8(o);
        }
    }
}

// The row stays PACKED (bf16 pairs, MAXV uint4 per lane) and is unpacked on the fly by each pass: half the registers of an fp32
// copy, so more rows are in flight per SM (the fp32 form ran at 0.49 of the HBM copy bandwidth with 16 warps per SM).  The
// arithmetic (order of the sums, the two-pass variance, affine then modulation) is unchanged.
template <int MAXV>
__global__ void __launch_bounds__(256, MAXV <= 12 ? 3 : 2)
adaln_modulate_kernel(const bf16* __restrict__ x, bf16* __restrict__ out, const bf16* __restrict__ ln_w,
                      const bf16* __restrict__ ln_b, const float* __restrict__ mod, int mod_stride, int shift_off_text,
                      int scale_off_text, int shift_off_other, int scale_off_other, long long rows, int S, int D,
                      int text_len, float eps) {
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int b = int(row / S);
    const int s = int(row - (long long)b * S);
    const int nvec = D / 8;
    uint4 raw[MAXV];
    const uint4* xr = reinterpret_cast<const uint4*>(x + row * D);
#pragma unroll
    for (int i = 0; i < MAXV; ++i) raw[i] = (i * 32 + lane < nvec) ? __ldg(xr + i * 32 + lane) : make_uint4(0u, 0u, 0u, 0u);
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
        float f[8];
        unpack8(raw[i], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) sum += f[j];
    }
    const float mean = warp_sum(sum) / float(D);
    // (the unpacked values must not be kept alive from one pass to the next — common-subexpression elimination would turn the
    // packed row back into an fp32 copy: "launder" the registers between passes)
#pragma unroll
    for (int i = 0; i < MAXV; ++i) asm volatile("" : "+r"(raw[i].x), "+r"(raw[i].y), "+r"(raw[i].z), "+r"(raw[i].w));
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
        if (i * 32 + lane < nvec) {
            float f[8];
            unpack8(raw[i], f);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float d = f[j] - mean;
                q += d * d;
            }
        }
    }
    const float rstd = rsqrtf(warp_sum(q) / float(D) + eps);
#pragma unroll
    for (int i = 0; i < MAXV; ++i) asm volatile("" : "+r"(raw[i].x), "+r"(raw[i].y), "+r"(raw[i].z), "+r"(raw[i].w));
    const float* m = mod + (long long)b * mod_stride;
    const bool text = s < text_len;
    const float* shift = m + (text ? shift_off_text : shift_off_other);
    const float* scale = m + (text ? scale_off_text : scale_off_other);
    uint4* orow = reinterpret_cast<uint4*>(out + row * D);
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
        const int vi = i * 32 + lane;
        if (vi < nvec) {
            float f[8], wf[8], bfv[8];
            unpack8(raw[i], f);
            unpack8(__ldg(reinterpret_cast<const uint4*>(ln_w) + vi), wf);
            unpack8(__ldg(reinterpret_cast<const uint4*>(ln_b) + vi), bfv);
            const float4 s0 = __ldg(reinterpret_cast<const float4*>(scale) + 2 * vi);
            const float4 s1 = __ldg(reinterpret_cast<const float4*>(scale) + 2 * vi + 1);
            const float4 h0 = __ldg(reinterpret_cast<const float4*>(shift) + 2 * vi);
            const float4 h1 = __ldg(reinterpret_cast<const float4*>(shift) + 2 * vi + 1);
            const float sc[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
            const float sh[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
            float o[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float y = (f[j] - mean) * rstd * wf[j] + bfv[j];
                o[j] = y * (1.0f + sc[j]) + sh[j];
            }
            orow[vi] = pack8(o);
        }
        asm volatile("" ::: "memory");   // keep the parameter loads of later vectors from being hoisted (register pressure -> spills)
    }
}

// Two rows per warp (round 2).  Per 16-byte vector of a row the single-row kernel also fetches 16 B of ln_w, 16 B of ln_b, 32 B of scale
// and 32 B of shift through L1: 96 B of parameter reads per 16 B of activations, i.e. ~155 B/clk/SM at the HBM rate — more than
// the 128 B/clk an SM's L1 delivers, which is what held the kernel at 0.59 of the copy bandwidth.  Two consecutive rows share one
// fetch of every parameter vector (rows of one batch and one segment share scale / shift too; the pair that straddles a
// text | other or batch boundary fetches its own).  Each row's arithmetic is exactly the single-row kernel's (bit-identical results).
// Statistics, affine and modulation of a PAIR of rows held packed in registers.  All arithmetic on fp32 PAIRS (fma/add/mul.rn.f32x2: the
// two bf16 halves of a word are one operand): ~8.8 instead of ~12.5 instructions per element (0.627 -> 0.65 of the copy bandwidth).
// Per element the operations and their order are the scalar kernel's ((x - mean) * rstd * w + b, then * (1 + scale) + shift, two-pass
// variance); only the row sums are accumulated in two interleaved partial sums (even / odd elements) instead of one.
// Measured and dropped (round 2): a persistent form whose row pairs are prefetched by cp.async.bulk into per-warp double buffers — 8
// warps per SM with 96 KB always in flight ran at 0.47 - 0.48 (bit-identical): 192 KB of row buffers leave ~64 KB of L1 and the
// per-vector parameter reads (ln_w, ln_b, scale, shift: 96 B per 16 B of activations) then come from L2.  The kernel is bound by its
// dependent instruction stream at ~0.1 IPC per warp, i.e. by warps per SM, not by bytes in flight.
template <int MAXV>
__device__ __forceinline__ void adaln_pair(uint4 (&ra)[MAXV], uint4 (&rb)[MAXV], bool has1, long long row0, int lane, int nvec,
                                           bf16* __restrict__ out, const bf16* __restrict__ ln_w, const bf16* __restrict__ ln_b,
                                           const float* __restrict__ mod, int mod_stride, int shift_off_text, int scale_off_text,
                                           int shift_off_other, int scale_off_other, int S, int D, int text_len, float eps) {
    auto word = [](const uint4& u, int k) -> uint32_t { return k == 0 ? u.x : k == 1 ? u.y : k == 2 ? u.z : u.w; };
    auto hsum = [](uint64_t v) -> float { float a, b; unpack2(v, a, b); return a + b; };
    uint64_t sa2 = 0ull, sb2 = 0ull;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            sa2 = fadd2(sa2, bf16x2_to_f32x2(word(ra[i], k)));
            sb2 = fadd2(sb2, bf16x2_to_f32x2(word(rb[i], k)));
        }
    }
    const float mean_a = warp_sum(hsum(sa2)) / float(D), mean_b = warp_sum(hsum(sb2)) / float(D);
    // (the unpacked values must not be kept alive from one pass to the next — common-subexpression elimination would turn the packed
    // rows back into fp32 copies: "launder" the registers between passes)
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
        asm volatile("" : "+r"(ra[i].x), "+r"(ra[i].y), "+r"(ra[i].z), "+r"(ra[i].w));
        asm volatile("" : "+r"(rb[i].x), "+r"(rb[i].y), "+r"(rb[i].z), "+r"(rb[i].w));
    }
    const uint64_t nma = pack2(-mean_a, -mean_a), nmb = pack2(-mean_b, -mean_b);
    uint64_t qa2 = 0ull, qb2 = 0ull;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
        if (i * 32 + lane < nvec) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint64_t da = fadd2(bf16x2_to_f32x2(word(ra[i], k)), nma), db = fadd2(bf16x2_to_f32x2(word(rb[i], k)), nmb);
                qa2 = ffma2(da, da, qa2);
                qb2 = ffma2(db, db, qb2);
            }
        }
    }
    const float rstd_a = rsqrtf(warp_sum(hsum(qa2)) / float(D) + eps), rstd_b = rsqrtf(warp_sum(hsum(qb2)) / float(D) + eps);
    // (the unpacked values must not be kept alive from one pass to the next — common-subexpression elimination would turn the packed
    // rows back into fp32 copies: "launder" the registers between passes)
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
        asm volatile("" : "+r"(ra[i].x), "+r"(ra[i].y), "+r"(ra[i].z), "+r"(ra[i].w));
        asm volatile("" : "+r"(rb[i].x), "+r"(rb[i].y), "+r"(rb[i].z), "+r"(rb[i].w));
    }
    const uint64_t rsa = pack2(rstd_a, rstd_a), rsb = pack2(rstd_b, rstd_b), one2 = pack2(1.0f, 1.0f);
    const int ba = int(row0 / S), sa_ = int(row0 - (long long)ba * S);
    const long long row1 = has1 ? row0 + 1 : row0;
    const int bb = int(row1 / S), sb_ = int(row1 - (long long)bb * S);
    const float* ma = mod + (long long)ba * mod_stride;
    const float* mb = mod + (long long)bb * mod_stride;
    const float* shift_a = ma + (sa_ < text_len ? shift_off_text : shift_off_other);
    const float* scale_a = ma + (sa_ < text_len ? scale_off_text : scale_off_other);
    const float* shift_b = mb + (sb_ < text_len ? shift_off_text : shift_off_other);
    const float* scale_b = mb + (sb_ < text_len ? scale_off_text : scale_off_other);
    const bool same = (shift_a == shift_b) && (scale_a == scale_b);      // warp-uniform
    uint4* oa = reinterpret_cast<uint4*>(out + row0 * D);
    uint4* ob = oa + nvec;
    auto pack_pair = [](uint64_t v) -> uint32_t { float a, b; unpack2(v, a, b); return pack_bf16x2(a, b); };
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
        const int vi = i * 32 + lane;
        if (vi < nvec) {
            const uint4 w4 = __ldg(reinterpret_cast<const uint4*>(ln_w) + vi), b4 = __ldg(reinterpret_cast<const uint4*>(ln_b) + vi);
            // 8 fp32 of scale and of shift = 4 pairs each, already in element-pair order
            const ulonglong2 sc01 = __ldg(reinterpret_cast<const ulonglong2*>(scale_a) + 2 * vi), sc23 = __ldg(reinterpret_cast<const ulonglong2*>(scale_a) + 2 * vi + 1);
            const ulonglong2 sh01 = __ldg(reinterpret_cast<const ulonglong2*>(shift_a) + 2 * vi), sh23 = __ldg(reinterpret_cast<const ulonglong2*>(shift_a) + 2 * vi + 1);
            if (has1 && !same) {      // the pair straddles a text | other or a batch boundary: row b has its own modulation vectors
                const ulonglong2 tc01 = __ldg(reinterpret_cast<const ulonglong2*>(scale_b) + 2 * vi), tc23 = __ldg(reinterpret_cast<const ulonglong2*>(scale_b) + 2 * vi + 1);
                const ulonglong2 th01 = __ldg(reinterpret_cast<const ulonglong2*>(shift_b) + 2 * vi), th23 = __ldg(reinterpret_cast<const ulonglong2*>(shift_b) + 2 * vi + 1);
                uint32_t oa4[4], ob4[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint64_t w2 = bf16x2_to_f32x2(word(w4, k)), b2 = bf16x2_to_f32x2(word(b4, k));
                    const uint64_t sca = fadd2(k == 0 ? sc01.x : k == 1 ? sc01.y : k == 2 ? sc23.x : sc23.y, one2);
                    const uint64_t sha = k == 0 ? sh01.x : k == 1 ? sh01.y : k == 2 ? sh23.x : sh23.y;
                    const uint64_t scb = fadd2(k == 0 ? tc01.x : k == 1 ? tc01.y : k == 2 ? tc23.x : tc23.y, one2);
                    const uint64_t shb = k == 0 ? th01.x : k == 1 ? th01.y : k == 2 ? th23.x : th23.y;
                    const uint64_t ta = fmul2(fadd2(bf16x2_to_f32x2(word(ra[i], k)), nma), rsa);
                    const uint64_t tb = fmul2(fadd2(bf16x2_to_f32x2(word(rb[i], k)), nmb), rsb);
                    oa4[k] = pack_pair(ffma2(ffma2(ta, w2, b2), sca, sha));
                    ob4[k] = pack_pair(ffma2(ffma2(tb, w2, b2), scb, shb));
                }
                oa[vi] = make_uint4(oa4[0], oa4[1], oa4[2], oa4[3]);
                ob[vi] = make_uint4(ob4[0], ob4[1], ob4[2], ob4[3]);
            } else {
                uint32_t oa4[4], ob4[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint64_t w2 = bf16x2_to_f32x2(word(w4, k)), b2 = bf16x2_to_f32x2(word(b4, k));
                    const uint64_t sc = fadd2(k == 0 ? sc01.x : k == 1 ? sc01.y : k == 2 ? sc23.x : sc23.y, one2);
                    const uint64_t sh = k == 0 ? sh01.x : k == 1 ? sh01.y : k == 2 ? sh23.x : sh23.y;
                    const uint64_t ta = fmul2(fadd2(bf16x2_to_f32x2(word(ra[i], k)), nma), rsa);
                    const uint64_t tb = fmul2(fadd2(bf16x2_to_f32x2(word(rb[i], k)), nmb), rsb);
                    oa4[k] = pack_pair(ffma2(ffma2(ta, w2, b2), sc, sh));
                    ob4[k] = pack_pair(ffma2(ffma2(tb, w2, b2), sc, sh));
                }
                oa[vi] = make_uint4(oa4[0], oa4[1], oa4[2], oa4[3]);
                if (has1) ob[vi] = make_uint4(ob4[0], ob4[1], ob4[2], ob4[3]);
            }
        }
        asm volatile("" ::: "memory");   // keep the parameter loads of later vectors from being hoisted (register pressure -> spills)
    }
}

template <int MAXV>
__global__ void __launch_bounds__(128, 3)      // 170 registers: two packed rows of up to 4096 elements without spills; 12 warps = 24 rows per SM
adaln_modulate2_kernel(const bf16* __restrict__ x, bf16* __restrict__ out, const bf16* __restrict__ ln_w,
                       const bf16* __restrict__ ln_b, const float* __restrict__ mod, int mod_stride, int shift_off_text,
                       int scale_off_text, int shift_off_other, int scale_off_other, long long rows, int S, int D,
                       int text_len, float eps) {
    const int lane = threadIdx.x & 31;
    const long long row0 = ((long long)blockIdx.x * 4 + (threadIdx.x >> 5)) * 2;
    if (row0 >= rows) return;
    const bool has1 = row0 + 1 < rows;
    const int nvec = D / 8;
    uint4 ra[MAXV], rb[MAXV];
    const uint4* xa = reinterpret_cast<const uint4*>(x + row0 * D);
    const uint4* xb = xa + nvec;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
        const bool in = i * 32 + lane < nvec;
        ra[i] = in ? __ldg(xa + i * 32 + lane) : make_uint4(0u, 0u, 0u, 0u);
        rb[i] = (in && has1) ? __ldg(xb + i * 32 + lane) : make_uint4(0u, 0u, 0u, 0u);
    }
    adaln_pair<MAXV>(ra, rb, has1, row0, lane, nvec, out, ln_w, ln_b, mod, mod_stride, shift_off_text, scale_off_text, shift_off_other,
                     scale_off_other, S, D, text_len, eps);
}

template <int MAXV>
__global__ void __launch_bounds__(256)
final_norm_kernel(const bf16* __restrict__ x, bf16* __restrict__ out, const bf16* __restrict__ w1, const bf16* __restrict__ b1,
                  const bf16* __restrict__ w2, const bf16* __restrict__ b2, const float* __restrict__ mod, int mod_stride,
                  int shift_off, int scale_off, int B, int S, int row0, int D, float eps) {
    const int lane = threadIdx.x & 31;
    const int R = S - row0;
    const long long orow = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (orow >= (long long)B * R) return;
    const int b = int(orow / R);
    const int s = row0 + int(orow - (long long)b * R);
    const int nvec = D / 8;
    RowRegs<MAXV> r;
    row_load<MAXV>(x + ((long long)b * S + s) * D, nvec, lane, r);
    float mean, rstd;
    row_stats<MAXV>(r, nvec, lane, D, eps, mean, rstd);
    row_affine<MAXV>(r, nvec, lane, mean, rstd, w1, b1);
    // the reference materialises norm_final's output in bf16 before norm_out sees it (cogvideox_transformer_3d.py:537-542)
#pragma unroll
    for (int i = 0; i < MAXV; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) r.v[i][j] = __bfloat162float(__float2bfloat16(r.v[i][j]));
    row_stats<MAXV>(r, nvec, lane, D, eps, mean, rstd);
    row_affine<MAXV>(r, nvec, lane, mean, rstd, w2, b2);
    const float* m = mod + (long long)b * mod_stride;
    row_modulate_store<MAXV>(r, nvec, lane, m + shift_off, m + scale_off, out + orow * D);
}

// ------------------------------------------------------------------------------------------ q/k LayerNorm + RoPE
// 8 lanes cooperate on one 64-wide head vector (16 bytes each); a 256-thread block walks the 2*H head vectors of one
// token row, so the row's cos/sin slice is loaded once per thread and reused for every head.
__global__ void __launch_bounds__(256)
qk_norm_rope_kernel(bf16* __restrict__ qkv, const bf16* __restrict__ nq_w, const bf16* __restrict__ nq_b,
                    const bf16* __restrict__ nk_w, const bf16* __restrict__ nk_b, const float* __restrict__ cos_t,
                    const float* __restrict__ sin_t, int S, int H, int text_len, float eps) {
    const long long row = blockIdx.x;  // b*S + s
    const int s = int(row % S);
    const int l8 = threadIdx.x & 7;
    const int slot = threadIdx.x >> 3;  // 0..31
    const bool rope = (cos_t != nullptr) && (s >= text_len);
    float c[8], sn[8];
    if (rope) {
        const float4* cp = reinterpret_cast<const float4*>(cos_t + (long long)(s - text_len) * 64 + l8 * 8);
        const float4* sp = reinterpret_cast<const float4*>(sin_t + (long long)(s - text_len) * 64 + l8 * 8);
        const float4 c0 = __ldg(cp), c1 = __ldg(cp + 1), s0 = __ldg(sp), s1 = __ldg(sp + 1);
        c[0] = c0.x; c[1] = c0.y; c[2] = c0.z; c[3] = c0.w; c[4] = c1.x; c[5] = c1.y; c[6] = c1.z; c[7] = c1.w;
        sn[0] = s0.x; sn[1] = s0.y; sn[2] = s0.z; sn[3] = s0.w; sn[4] = s1.x; sn[5] = s1.y; sn[6] = s1.z; sn[7] = s1.w;
    }
    float wq[8], bq[8], wk[8], bk[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(nq_w) + l8), wq);
    unpack8(__ldg(reinterpret_cast<const uint4*>(nq_b) + l8), bq);
    unpack8(__ldg(reinterpret_cast<const uint4*>(nk_w) + l8), wk);
    unpack8(__ldg(reinterpret_cast<const uint4*>(nk_b) + l8), bk);
    bf16* base = qkv + row * (long long)(3 * H * 64);
    for (int hv = slot; hv < 2 * H; hv += 32) {  // hv < H: q heads, else k heads (contiguous in the row)
        uint4* ptr = reinterpret_cast<uint4*>(base + hv * 64) + l8;
        float v[8];
        unpack8(*ptr, v);
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) sum += v[j];
        sum += __shfl_xor_sync(0xffffffffu, sum, 1);
        sum += __shfl_xor_sync(0xffffffffu, sum, 2);
        sum += __shfl_xor_sync(0xffffffffu, sum, 4);
        const float mean = sum * (1.0f / 64.0f);
        float sq = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float d = v[j] - mean;
            sq = fmaf(d, d, sq);
        }
        sq += __shfl_xor_sync(0xffffffffu, sq, 1);
        sq += __shfl_xor_sync(0xffffffffu, sq, 2);
        sq += __shfl_xor_sync(0xffffffffu, sq, 4);
        const float rstd = rsqrtf(sq * (1.0f / 64.0f) + eps);
        const bool is_q = hv < H;
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = fmaf((v[j] - mean) * rstd, is_q ? wq[j] : wk[j], is_q ? bq[j] : bk[j]);
        if (rope) {
            // the reference rounds the LayerNorm output to bf16 before the fp32 rotation (attention_processor.py:2060-2066)
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = __bfloat162float(__float2bfloat16(v[j]));
            float o[8];
#pragma unroll
            for (int j = 0; j < 8; j += 2) {
                // explicit FMAs: head_norm_rope in gemm_tcgen05.cu (the fused epilogue) must round identically
                o[j] = fmaf(v[j], c[j], -(v[j + 1] * sn[j]));
                o[j + 1] = fmaf(v[j + 1], c[j + 1], v[j] * sn[j + 1]);
            }
            *ptr = pack8(o);
        } else {
            *ptr = pack8(v);
        }
    }
}

// ------------------------------------------------------------------------------------------ small-batch linear
// One warp per output feature n; x rows (B <= 8) stay in registers / L1.  out = beta*out + alpha*(w.f(x) + bias).
__global__ void __launch_bounds__(256)
small_linear_kernel(const float* __restrict__ x, long long ldx, const bf16* __restrict__ w, long long ldw,
                    const bf16* __restrict__ bias, float* __restrict__ out, long long ldo, int B, int N, int K, int act_in,
                    float alpha, float beta, int round_bf16) {
    const int lane = threadIdx.x & 31;
    const int n = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (n >= N) return;
    float acc[8];
#pragma unroll
    for (int b = 0; b < 8; ++b) acc[b] = 0.f;
    const bf16* wr = w + (long long)n * ldw;
    for (int k0 = lane * 8; k0 < K; k0 += 256) {
        float wf[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(wr + k0)), wf);
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            if (b < B) {
                const float4 x0 = __ldg(reinterpret_cast<const float4*>(x + b * ldx + k0));
                const float4 x1 = __ldg(reinterpret_cast<const float4*>(x + b * ldx + k0 + 4));
                float xv[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
                if (act_in == 1) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) xv[j] = xv[j] / (1.0f + __expf(-xv[j]));
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[b] = fmaf(wf[j], xv[j], acc[b]);
            }
        }
    }
#pragma unroll
    for (int b = 0; b < 8; ++b) acc[b] = warp_sum(acc[b]);
    if (lane == 0) {
        const float bv = bias ? __bfloat162float(bias[n]) : 0.f;
        for (int b = 0; b < B; ++b) {
            float r = alpha * (acc[b] + bv);
            if (beta != 0.f) r += beta * out[b * ldo + n];
            if (round_bf16) r = __bfloat162float(__float2bfloat16(r));
            out[b * ldo + n] = r;
        }
    }
}

__global__ void timestep_sinusoid_kernel(const float* __restrict__ t, const float* __restrict__ freqs, float* __restrict__ out,
                                         int B, int D, int round_bf16) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int half = D / 2;
    if (idx >= B * half) return;
    const int b = idx / half, i = idx - b * half;
    const float arg = __fmul_rn(t[b], freqs[i]);   // freqs = exp(-ln(10000) * i / half), tabulated by the host in fp32
    float c = cosf(arg), s = sinf(arg);
    if (round_bf16) {
        c = __bfloat162float(__float2bfloat16(c));
        s = __bfloat162float(__float2bfloat16(s));
    }
    out[b * D + i] = c;          // flip_sin_to_cos: cos first
    out[b * D + half + i] = s;
}

// The same arithmetic for MANY independent problems in one launch (blockIdx.y = problem): the 2 x layers + 1 AdaLN modulation linears of a
// step (and their LoRA pairs) are 255 launches of ~15 us each when issued one by one — each streams 19 MB of weights at a fifth of the HBM
// rate — and one launch per dependency stage this way.  A problem whose x is null reads the launch's default input (the time embedding).
__global__ void __launch_bounds__(256)
small_linear_batch_kernel(const s2v_small_linear_desc* __restrict__ descs, const float* __restrict__ x_default, long long ldx_default, int B) {
    const s2v_small_linear_desc d = descs[blockIdx.y];
    const int lane = threadIdx.x & 31;
    const int n = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (n >= d.N) return;
    const float* x = d.x ? d.x : x_default;
    const long long ldx = d.x ? d.ldx : ldx_default;
    float acc[8];
#pragma unroll
    for (int b = 0; b < 8; ++b) acc[b] = 0.f;
    const bf16* wr = static_cast<const bf16*>(d.w) + (long long)n * d.ldw;
    for (int k0 = lane * 8; k0 < d.K; k0 += 256) {
        float wf[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(wr + k0)), wf);
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            if (b < B) {
                const float4 x0 = __ldg(reinterpret_cast<const float4*>(x + b * ldx + k0));
                const float4 x1 = __ldg(reinterpret_cast<const float4*>(x + b * ldx + k0 + 4));
                float xv[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
                if (d.act_in == 1) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) xv[j] = xv[j] / (1.0f + __expf(-xv[j]));
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[b] = fmaf(wf[j], xv[j], acc[b]);
            }
        }
    }
#pragma unroll
    for (int b = 0; b < 8; ++b) acc[b] = warp_sum(acc[b]);
    if (lane == 0) {
        const float bv = d.bias ? __bfloat162float(static_cast<const bf16*>(d.bias)[n]) : 0.f;
        for (int b = 0; b < B; ++b) {
            float r = d.alpha * (acc[b] + bv);
            if (d.beta != 0.f) r += d.beta * d.out[b * d.ldo + n];
            if (d.round_bf16) r = __bfloat162float(__float2bfloat16(r));
            d.out[b * d.ldo + n] = r;
        }
    }
}

// ------------------------------------------------------------------------------------------ patchify / unpatchify
__global__ void patchify_kernel(const bf16* __restrict__ lat, bf16* __restrict__ rows, int NB, int C, int H, int W, int p) {
    const int hp = H / p, wp = W / p, K = C * p * p;
    const long long total = (long long)NB * hp * wp * K;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int k = int(idx % K);
        const long long tok = idx / K;
        const int j = int(tok % wp), i = int((tok / wp) % hp);
        const long long nb = tok / ((long long)wp * hp);
        const int c = k / (p * p), dy = (k / p) % p, dx = k % p;
        rows[idx] = lat[((nb * C + c) * H + (i * p + dy)) * W + (j * p + dx)];
    }
}

__global__ void unpatchify_kernel(const bf16* __restrict__ tok, bf16* __restrict__ lat, int NB, int C, int H, int W, int p) {
    const int hp = H / p, wp = W / p, K = C * p * p;
    const long long total = (long long)NB * C * H * W;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int x = int(idx % W), y = int((idx / W) % H);
        const int c = int((idx / ((long long)W * H)) % C);
        const long long nb = idx / ((long long)W * H * C);
        const long long t = (nb * hp + y / p) * wp + x / p;
        lat[idx] = tok[t * K + c * p * p + (y % p) * p + (x % p)];
    }
}

__global__ void add_rows_kernel(bf16* __restrict__ dst, const bf16* __restrict__ table, int B, int S, int D, int row0, int R) {
    const int nvec = D / 8;
    const long long total = (long long)B * R * nvec;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int v = int(idx % nvec);
        const long long br = idx / nvec;
        const int r = int(br % R);
        const long long b = br / R;
        uint4* d = reinterpret_cast<uint4*>(dst + (b * S + row0 + r) * (long long)D) + v;
        float a[8], t[8];
        unpack8(*d, a);
        unpack8(__ldg(reinterpret_cast<const uint4*>(table + (long long)r * D) + v), t);
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] += t[j];
        *d = pack8(a);
    }
}

// ------------------------------------------------------------------------------------------ CFG + DDIM (bit-exact)
__device__ __forceinline__ float round_bf16f(float x) { return __bfloat162float(__float2bfloat16(x)); }

__global__ void cfg_ddim_kernel(const bf16* __restrict__ noise, const bf16* __restrict__ lat, bf16* __restrict__ lat_out,
                                float* __restrict__ x0_out, long long n, float g, float sa, float sb, float ac, float bc) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float u = __bfloat162float(noise[i]);
        const float t = __bfloat162float(noise[n + i]);
        // noise_pred = u + g * (t - u)     — three separately rounded fp32 ops (custom_cogvideox_pipe.py:277-279)
        const float v = __fadd_rn(u, __fmul_rn(g, __fsub_rn(t, u)));
        const float x = __bfloat162float(lat[i]);
        // x0 = bf16(sa * x) - sb * v       (scheduling_ddim_cogvideox.py:383; scalar*bf16 tensor rounds to bf16)
        const float x0 = __fsub_rn(round_bf16f(__fmul_rn(sa, x)), __fmul_rn(sb, v));
        // prev = bf16(a * x) + b * x0      (:391-394), then .to(bf16) in the pipe (:296)
        const float prev = __fadd_rn(round_bf16f(__fmul_rn(ac, x)), __fmul_rn(bc, x0));
        lat_out[i] = __float2bfloat16(prev);
        if (x0_out) x0_out[i] = x0;
    }
}

// plain DDIM step on an fp32 model output (the scheduler.step() surface): prev, x0 in fp32
__global__ void ddim_step_kernel(const float* __restrict__ v_in, const bf16* __restrict__ lat, float* __restrict__ prev_out,
                                 float* __restrict__ x0_out, long long n, float sa, float sb, float ac, float bc) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float v = v_in[i];
        const float x = __bfloat162float(lat[i]);
        const float x0 = __fsub_rn(round_bf16f(__fmul_rn(sa, x)), __fmul_rn(sb, v));
        prev_out[i] = __fadd_rn(round_bf16f(__fmul_rn(ac, x)), __fmul_rn(bc, x0));
        if (x0_out) x0_out[i] = x0;
    }
}

// DPM-Solver++ (SDE, multistep) update of CogVideoXDPMScheduler.step (scheduling_dpm_cogvideox.py:383-439) with the
// reference's rounding points on CUDA: (fp64 scalar) * (bf16 tensor) is an fp32 product rounded to bf16, everything that
// touches an fp32 tensor stays fp32, left-to-right evaluation, no FMA contraction.
//   x0   = bf16(sa*x) - sb*v
//   d    = second_order ? m2*x0 - m3*old_x0 : x0
//   prev = (bf16(m0*x) - m1*d) + bf16(mn*noise)
// CFG = true: v = u + g*(t - u) from a bf16 [2, n] model output (uncond first), result rounded to bf16 (the pipe's .to());
// CFG = false: v is the fp32 model output and prev stays fp32 (the scheduler.step surface).
template <bool CFG>
__global__ void dpm_step_kernel(const void* __restrict__ v_in, const bf16* __restrict__ lat, const float* __restrict__ old_x0,
                                const bf16* __restrict__ noise, void* __restrict__ prev_out, float* __restrict__ x0_out, long long n,
                                float g, float sa, float sb, float m0, float m1, float m2, float m3, float mn) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float v;
        if (CFG) {
            const bf16* np = static_cast<const bf16*>(v_in);
            const float u = __bfloat162float(np[i]), t = __bfloat162float(np[n + i]);
            v = __fadd_rn(u, __fmul_rn(g, __fsub_rn(t, u)));
        } else {
            v = static_cast<const float*>(v_in)[i];
        }
        const float x = __bfloat162float(lat[i]);
        const float x0 = __fsub_rn(round_bf16f(__fmul_rn(sa, x)), __fmul_rn(sb, v));
        const float d = old_x0 ? __fsub_rn(__fmul_rn(m2, x0), __fmul_rn(m3, old_x0[i])) : x0;
        const float prev = __fadd_rn(__fsub_rn(round_bf16f(__fmul_rn(m0, x)), __fmul_rn(m1, d)),
                                     round_bf16f(__fmul_rn(mn, __bfloat162float(noise[i]))));
        if (CFG) static_cast<bf16*>(prev_out)[i] = __float2bfloat16(prev);
        else static_cast<float*>(prev_out)[i] = prev;
        x0_out[i] = x0;
    }
}

template <typename F>
static int dispatch_maxv(int D, F&& f) {
    const int nvec = D / 8;
    if (nvec <= 32) return f(std::integral_constant<int, 1>{});
    if (nvec <= 128) return f(std::integral_constant<int, 4>{});
    if (nvec <= 256) return f(std::integral_constant<int, 8>{});
    if (nvec <= 384) return f(std::integral_constant<int, 12>{});
    if (nvec <= 512) return f(std::integral_constant<int, 16>{});
    return set_error(S2V_E_UNSUPPORTED, "LayerNorm width > 4096 not supported");
}

}  // namespace s2v

using namespace s2v;

#define S2V_PROLOGUE()                                              \
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);      \
    {                                                               \
        int rc_ = ensure_device();                                  \
        if (rc_) return rc_;                                        \
    }

extern "C" int s2v_adaln_modulate(const void* x, void* out, const void* ln_w, const void* ln_b, const float* mod,
                                  int32_t mod_stride, int32_t shift_off_text, int32_t scale_off_text,
                                  int32_t shift_off_other, int32_t scale_off_other, int32_t B, int32_t S, int32_t D,
                                  int32_t text_len, float eps, void* stream_) {
    if (!x || !out || !ln_w || !ln_b || !mod) return set_error(S2V_E_BADARG, "s2v_adaln_modulate: null pointer");
    if (B <= 0 || S <= 0 || D <= 0 || (D % 8)) return set_error(S2V_E_BADARG, "s2v_adaln_modulate: bad shape (D % 8 != 0?)");
    if ((shift_off_text | scale_off_text | shift_off_other | scale_off_other | mod_stride) % 4)
        return set_error(S2V_E_BADARG, "s2v_adaln_modulate: modulation offsets must be multiples of 4 floats");
    S2V_PROLOGUE();
    const long long rows = (long long)B * S;
    // two rows per warp (see adaln_modulate2_kernel); S2V_ADALN_ROWS=1 keeps the single-row kernel (A/B, bit-identical)
    static const bool two_rows = [] { const char* e = getenv("S2V_ADALN_ROWS"); return !(e && e[0] == '1'); }();
    if (two_rows && D >= 1024) {
        const unsigned grid2 = (unsigned)((rows + 7) / 8);
        return dispatch_maxv(D, [&](auto mv) {
            adaln_modulate2_kernel<decltype(mv)::value><<<grid2, 128, 0, stream>>>(
                static_cast<const bf16*>(x), static_cast<bf16*>(out), static_cast<const bf16*>(ln_w),
                static_cast<const bf16*>(ln_b), mod, mod_stride, shift_off_text, scale_off_text, shift_off_other,
                scale_off_other, rows, S, D, text_len, eps);
            return check_launch("adaln_modulate2_kernel");
        });
    }
    const unsigned grid = (unsigned)((rows + 7) / 8);
    return dispatch_maxv(D, [&](auto mv) {
        adaln_modulate_kernel<decltype(mv)::value><<<grid, 256, 0, stream>>>(
            static_cast<const bf16*>(x), static_cast<bf16*>(out), static_cast<const bf16*>(ln_w),
            static_cast<const bf16*>(ln_b), mod, mod_stride, shift_off_text, scale_off_text, shift_off_other,
            scale_off_other, rows, S, D, text_len, eps);
        return check_launch("adaln_modulate_kernel");
    });
}

extern "C" int s2v_final_norm(const void* x, void* out, const void* ln1_w, const void* ln1_b, const void* ln2_w,
                              const void* ln2_b, const float* mod, int32_t mod_stride, int32_t shift_off,
                              int32_t scale_off, int32_t B, int32_t S, int32_t row0, int32_t D, float eps, void* stream_) {
    if (!x || !out || !ln1_w || !ln1_b || !ln2_w || !ln2_b || !mod) return set_error(S2V_E_BADARG, "s2v_final_norm: null pointer");
    if (B <= 0 || S <= 0 || row0 < 0 || row0 >= S || (D % 8)) return set_error(S2V_E_BADARG, "s2v_final_norm: bad shape");
    S2V_PROLOGUE();
    const long long rows = (long long)B * (S - row0);
    const unsigned grid = (unsigned)((rows + 7) / 8);
    return dispatch_maxv(D, [&](auto mv) {
        final_norm_kernel<decltype(mv)::value><<<grid, 256, 0, stream>>>(
            static_cast<const bf16*>(x), static_cast<bf16*>(out), static_cast<const bf16*>(ln1_w),
            static_cast<const bf16*>(ln1_b), static_cast<const bf16*>(ln2_w), static_cast<const bf16*>(ln2_b), mod,
            mod_stride, shift_off, scale_off, B, S, row0, D, eps);
        return check_launch("final_norm_kernel");
    });
}

extern "C" int s2v_qk_norm_rope(void* qkv, const void* nq_w, const void* nq_b, const void* nk_w, const void* nk_b,
                                const float* cos, const float* sin, int32_t B, int32_t S, int32_t H, int32_t text_len,
                                float eps, void* stream_) {
    if (!qkv || !nq_w || !nq_b || !nk_w || !nk_b) return set_error(S2V_E_BADARG, "s2v_qk_norm_rope: null pointer");
    if ((cos == nullptr) != (sin == nullptr)) return set_error(S2V_E_BADARG, "s2v_qk_norm_rope: cos/sin must both be set or both NULL");
    if (B <= 0 || S <= 0 || H <= 0 || text_len < 0 || text_len > S) return set_error(S2V_E_BADARG, "s2v_qk_norm_rope: bad shape");
    S2V_PROLOGUE();
    qk_norm_rope_kernel<<<(unsigned)((long long)B * S), 256, 0, stream>>>(
        static_cast<bf16*>(qkv), static_cast<const bf16*>(nq_w), static_cast<const bf16*>(nq_b),
        static_cast<const bf16*>(nk_w), static_cast<const bf16*>(nk_b), cos, sin, S, H, text_len, eps);
    return check_launch("qk_norm_rope_kernel");
}

extern "C" int s2v_small_linear(const float* x, int64_t ldx, const void* w, int64_t ldw, const void* bias, float* out,
                                int64_t ldo, int32_t B, int32_t N, int32_t K, int32_t act_in, float alpha, float beta,
                                int32_t round_bf16, void* stream_) {
    if (!x || !w || !out) return set_error(S2V_E_BADARG, "s2v_small_linear: null pointer");
    if (B <= 0 || B > 8 || N <= 0 || K <= 0 || (K % 8) || (ldx % 4) || (ldw % 8))
        return set_error(S2V_E_UNSUPPORTED, "s2v_small_linear: need 1 <= B <= 8, K % 8 == 0, aligned leading dims");
    S2V_PROLOGUE();
    small_linear_kernel<<<(N + 7) / 8, 256, 0, stream>>>(x, ldx, static_cast<const bf16*>(w), ldw,
                                                         static_cast<const bf16*>(bias), out, ldo, B, N, K, act_in, alpha,
                                                         beta, round_bf16);
    return check_launch("small_linear_kernel");
}

extern "C" int s2v_small_linear_batch(const s2v_small_linear_desc* descs, int32_t count, int32_t max_n, const float* x_default,
                                      int64_t ldx_default, int32_t B, void* stream_) {
    if (!descs || count <= 0 || max_n <= 0) return set_error(S2V_E_BADARG, "s2v_small_linear_batch: empty batch");
    if (B <= 0 || B > 8 || count > 65535 || (ldx_default % 4)) return set_error(S2V_E_UNSUPPORTED, "s2v_small_linear_batch: need 1 <= B <= 8, count <= 65535");
    S2V_PROLOGUE();
    small_linear_batch_kernel<<<dim3((max_n + 7) / 8, count), 256, 0, stream>>>(descs, x_default, ldx_default, B);
    return check_launch("small_linear_batch_kernel");
}

extern "C" int s2v_timestep_sinusoid(const float* t, const float* freqs, float* out, int32_t B, int32_t D, int32_t round_bf16,
                                     void* stream_) {
    if (!t || !freqs || !out || B <= 0 || D <= 0 || (D % 2)) return set_error(S2V_E_BADARG, "s2v_timestep_sinusoid: bad argument");
    S2V_PROLOGUE();
    const int n = B * (D / 2);
    timestep_sinusoid_kernel<<<(n + 255) / 256, 256, 0, stream>>>(t, freqs, out, B, D, round_bf16);
    return check_launch("timestep_sinusoid_kernel");
}

static unsigned ew_grid(long long total) {
    long long g = (total + 255) / 256;
    const long long cap = (long long)sm_count() * 16;
    return (unsigned)(g < cap ? (g > 0 ? g : 1) : cap);
}

extern "C" int s2v_patchify(const void* latents, void* rows, int32_t NB, int32_t C, int32_t H, int32_t W, int32_t p, void* stream_) {
    if (!latents || !rows || NB <= 0 || C <= 0 || p <= 0 || (H % p) || (W % p)) return set_error(S2V_E_BADARG, "s2v_patchify: bad argument");
    S2V_PROLOGUE();
    patchify_kernel<<<ew_grid((long long)NB * C * H * W), 256, 0, stream>>>(static_cast<const bf16*>(latents),
                                                                           static_cast<bf16*>(rows), NB, C, H, W, p);
    return check_launch("patchify_kernel");
}

extern "C" int s2v_unpatchify(const void* tokens, void* latents, int32_t NB, int32_t C, int32_t H, int32_t W, int32_t p, void* stream_) {
    if (!tokens || !latents || NB <= 0 || C <= 0 || p <= 0 || (H % p) || (W % p)) return set_error(S2V_E_BADARG, "s2v_unpatchify: bad argument");
    S2V_PROLOGUE();
    unpatchify_kernel<<<ew_grid((long long)NB * C * H * W), 256, 0, stream>>>(static_cast<const bf16*>(tokens),
                                                                             static_cast<bf16*>(latents), NB, C, H, W, p);
    return check_launch("unpatchify_kernel");
}

extern "C" int s2v_add_rows(void* dst, const void* table, int32_t B, int32_t S, int32_t D, int32_t row0, int32_t R, void* stream_) {
    if (!dst || !table || B <= 0 || R <= 0 || row0 < 0 || row0 + R > S || (D % 8)) return set_error(S2V_E_BADARG, "s2v_add_rows: bad argument");
    S2V_PROLOGUE();
    add_rows_kernel<<<ew_grid((long long)B * R * (D / 8)), 256, 0, stream>>>(static_cast<bf16*>(dst),
                                                                            static_cast<const bf16*>(table), B, S, D, row0, R);
    return check_launch("add_rows_kernel");
}

extern "C" int s2v_cfg_ddim_step(const void* noise_pred, const void* latents, void* latents_out, float* x0_out,
                                 int64_t n_per_half, float guidance, float sqrt_alpha, float sqrt_beta, float a_coef,
                                 float b_coef, void* stream_) {
    if (!noise_pred || !latents || !latents_out || n_per_half <= 0) return set_error(S2V_E_BADARG, "s2v_cfg_ddim_step: bad argument");
    S2V_PROLOGUE();
    cfg_ddim_kernel<<<ew_grid(n_per_half), 256, 0, stream>>>(static_cast<const bf16*>(noise_pred),
                                                            static_cast<const bf16*>(latents), static_cast<bf16*>(latents_out),
                                                            x0_out, n_per_half, guidance, sqrt_alpha, sqrt_beta, a_coef, b_coef);
    return check_launch("cfg_ddim_kernel");
}

extern "C" int s2v_ddim_step(const float* model_output, const void* sample, float* prev_out, float* x0_out, int64_t n,
                             float sqrt_alpha, float sqrt_beta, float a_coef, float b_coef, void* stream_) {
    if (!model_output || !sample || !prev_out || n <= 0) return set_error(S2V_E_BADARG, "s2v_ddim_step: bad argument");
    S2V_PROLOGUE();
    ddim_step_kernel<<<ew_grid(n), 256, 0, stream>>>(model_output, static_cast<const bf16*>(sample), prev_out, x0_out, n,
                                                    sqrt_alpha, sqrt_beta, a_coef, b_coef);
    return check_launch("ddim_step_kernel");
}

extern "C" int s2v_dpm_step(const void* model_output, int32_t cfg_input, const void* sample, const float* old_x0, const void* noise,
                            void* prev_out, float* x0_out, int64_t n, float guidance, float sqrt_alpha, float sqrt_beta, float m0, float m1,
                            float m2, float m3, float m_noise, void* stream_) {
    if (!model_output || !sample || !noise || !prev_out || !x0_out) return set_error(S2V_E_BADARG, "s2v_dpm_step: null pointer");
    if (n <= 0) return set_error(S2V_E_BADARG, "s2v_dpm_step: empty problem");
    S2V_PROLOGUE();
    long long blocks = (n + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (cfg_input)
        dpm_step_kernel<true><<<(unsigned)blocks, 256, 0, stream>>>(model_output, static_cast<const bf16*>(sample), old_x0,
                                                                     static_cast<const bf16*>(noise), prev_out, x0_out, n, guidance, sqrt_alpha,
                                                                     sqrt_beta, m0, m1, m2, m3, m_noise);
    else
        dpm_step_kernel<false><<<(unsigned)blocks, 256, 0, stream>>>(model_output, static_cast<const bf16*>(sample), old_x0,
                                                                      static_cast<const bf16*>(noise), prev_out, x0_out, n, guidance, sqrt_alpha,
                                                                      sqrt_beta, m0, m1, m2, m3, m_noise);
    return check_launch("dpm_step_kernel");
}
