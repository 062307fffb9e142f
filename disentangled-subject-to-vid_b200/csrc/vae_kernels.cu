// HBM-bound kernels of the 3D causal VAE decoder.  Activations are padded channels-last VOLUMES
// [2 + T, Hp, Wp, C] bf16 (see include/s2v_b200.h); the convolutions themselves are s2v_conv_gemm (gemm_tcgen05.cu).
//   latent_rows / latent_im2col   latent tile [16, T, h, w] -> A operands of the 1x1x1 (conv_y | conv_b) GEMM and of conv_in
//   groupnorm_stats               GroupNorm(32) statistics over the T frames of a call: deterministic two-stage reduction
//   spatialnorm_silu              SiLU( GN(f) * conv_y(zq') + conv_b(zq') ) with zq' gathered by nearest-neighbour indices
//   upsample_nearest              CogVideoXUpsample3D's interpolate step, written straight into the padded conv input
//   volume_to_video               conv_out volume -> [3, T, H, W]
//   blend                         linear seam ramps of tiled_decode
// All loads/stores are 16-byte vectors along the channel dimension; fp32 arithmetic; grid sizes are multiples of 148.
#include "common.cuh"
#include "host_util.h"
#include "s2v_b200.h"

namespace s2v {

__device__ __forceinline__ void v_unpack8(const uint4& u, float (&f)[8]) {
    f[0] = bf16_lo(u.x); f[1] = bf16_hi(u.x); f[2] = bf16_lo(u.y); f[3] = bf16_hi(u.y);
    f[4] = bf16_lo(u.z); f[5] = bf16_hi(u.z); f[6] = bf16_lo(u.w); f[7] = bf16_hi(u.w);
}
__device__ __forceinline__ uint4 v_pack8(const float (&f)[8]) {
    uint4 u;
    u.x = pack_bf16x2(f[0], f[1]); u.y = pack_bf16x2(f[2], f[3]);
    u.z = pack_bf16x2(f[4], f[5]); u.w = pack_bf16x2(f[6], f[7]);
    return u;
}
__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

// ------------------------------------------------------------------------------------------------ latent -> GEMM operands
// z: one sample [C, Tz, hz, wz] bf16 (contiguous); the tile is rows [i0, i0+ht) x cols [j0, j0+wt), frames [f0, f0+T).
// rows_out [T*ht*wt, ldr] (ldr >= C, extra columns zero): channels-last rows of  scale * z  (the reference multiplies by
// 1/scaling_factor in bf16: pipeline_cogvideox.py:348).
__global__ void latent_rows_kernel(const bf16* __restrict__ z, bf16* __restrict__ rows, int C, int Tz, int hz, int wz, int f0,
                                   int T, int i0, int j0, int ht, int wt, int ldr, float scale) {
    const long long n = (long long)T * ht * wt * ldr;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % ldr);
        long long r = i / ldr;
        const int w = (int)(r % wt);
        r /= wt;
        const int h = (int)(r % ht);
        const int t = (int)(r / ht);
        float v = 0.f;
        if (c < C) v = __bfloat162float(z[(((long long)c * Tz + (f0 + t)) * hz + (i0 + h)) * wz + (j0 + w)]) * scale;
        rows[i] = __float2bfloat16_rn(v);
    }
}

// im2col of the causal 3x3x3 conv_in over the PADDED output grid: col[(t*Hp + hp)*Wp + wp, tap*C + c] with
// tap = (dt*3+dh)*3+dw reading frame max(f0 + t + dt - 2, 0) (the conv cache of conv_in is the latent itself; the
// first call replicates frame 0: autoencoder_kl_cogvideox.py:120-127) and zero spatial padding.
__global__ void latent_im2col_kernel(const bf16* __restrict__ z, bf16* __restrict__ col, int C, int Tz, int hz, int wz, int f0,
                                     int T, int i0, int j0, int ht, int wt, float scale) {
    const int Hp = ht + 2, Wp = wt + 2, K = 27 * C;
    const long long n = (long long)T * Hp * Wp * K;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(i % K);
        long long r = i / K;
        const int wp = (int)(r % Wp);
        r /= Wp;
        const int hp = (int)(r % Hp);
        const int t = (int)(r / Hp);
        const int tap = k / C, c = k - tap * C;
        const int dt = tap / 9, dh = (tap / 3) % 3, dw = tap % 3;
        const int h = hp - 1 + dh - 1, w = wp - 1 + dw - 1;
        int f = f0 + t + dt - 2;
        f = f < 0 ? 0 : f;
        float v = 0.f;
        if (h >= 0 && h < ht && w >= 0 && w < wt) v = __bfloat162float(z[(((long long)c * Tz + f) * hz + (i0 + h)) * wz + (j0 + w)]) * scale;
        col[i] = __float2bfloat16_rn(v);
    }
}

// ------------------------------------------------------------------------------------------------ GroupNorm statistics
// Stage 1: block b reduces lines [b*lpb, (b+1)*lpb) (a line = the W interior positions of one (t, h)) into per-channel
// sum / sum of squares.  blockDim = 256 and C/8 divides 256, so a thread always meets the same 8 channels.
__global__ void __launch_bounds__(256) gn_partial_kernel(const bf16* __restrict__ x, float* __restrict__ partial, int T, int H, int W,
                                                         int C, int lines_per_block) {
    const int Hp = H + 2, Wp = W + 2;
    const int vpp = C >> 3;                      // 16-byte vectors per position
    const int slot = threadIdx.x % vpp;
    const int pos0 = threadIdx.x / vpp, pstep = 256 / vpp;
    float s[8], q[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] = q[j] = 0.f;
    const int l0 = blockIdx.x * lines_per_block;
    const int l1 = min(l0 + lines_per_block, T * H);
    for (int l = l0; l < l1; ++l) {
        const int t = l / H, h = l - t * H;
        const uint4* line = reinterpret_cast<const uint4*>(x + ((long long)((t + 2) * Hp + h + 1) * Wp + 1) * C);
        // four independent 16-byte loads in flight per thread (the grid is only a few hundred blocks: one load per thread
        // measured 26 % of the HBM copy bandwidth); the accumulation order per thread stays p, p+pstep, p+2*pstep, ...
        for (int p = pos0; p < W; p += 4 * pstep) {
            uint4 u[4];
#pragma unroll
            for (int k = 0; k < 4; ++k)
                u[k] = (p + k * pstep < W) ? __ldg(line + (long long)(p + k * pstep) * vpp + slot) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float f[8];
                v_unpack8(u[k], f);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    s[j] += f[j];
                    q[j] += f[j] * f[j];
                }
            }
        }
    }
    __shared__ float red[256][17];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        red[threadIdx.x][j] = s[j];
        red[threadIdx.x][8 + j] = q[j];
    }
    __syncthreads();
    // thread (slot, j): sum over the pstep threads that share the slot, in a fixed order
    for (int o = threadIdx.x; o < vpp * 16; o += 256) {
        const int sl = o / 16, j = o % 16;
        float acc = 0.f;
        for (int k = 0; k < pstep; ++k) acc += red[k * vpp + sl][j];
        const int c = sl * 8 + (j & 7);
        partial[((long long)blockIdx.x * C + c) * 2 + (j >> 3)] = acc;
    }
}

// Stage 2: one GN_FIN_THREADS-thread block per group; stats[g] = (mean, rstd) over the group's channels and all stage-1 row blocks
// (up to a few thousand when they come out of a convolution epilogue).  Each thread sums a fixed strided subset in double, then a
// fixed-shape tree combines them (deterministic).
constexpr int GN_FIN_THREADS = 512;
__global__ void __launch_bounds__(GN_FIN_THREADS) gn_finalize_kernel(const float* __restrict__ partial, float* __restrict__ stats, int nblk, int C,
                                                                     int G, float count, float eps) {
    const int g = blockIdx.x, cpg = C / G, n = nblk * cpg;
    double s = 0.0, q = 0.0;
    for (int i = threadIdx.x; i < n; i += GN_FIN_THREADS) {
        const int b = i / cpg, c = g * cpg + (i - b * cpg);
        const float2 v = __ldg(reinterpret_cast<const float2*>(partial + ((long long)b * C + c) * 2));
        s += v.x;
        q += v.y;
    }
    __shared__ double rs[GN_FIN_THREADS], rq[GN_FIN_THREADS];
    rs[threadIdx.x] = s;
    rq[threadIdx.x] = q;
    __syncthreads();
    for (int o = GN_FIN_THREADS / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            rs[threadIdx.x] += rs[threadIdx.x + o];
            rq[threadIdx.x] += rq[threadIdx.x + o];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const double mean = rs[0] / count;
        double var = rq[0] / count - mean * mean;
        var = var < 0.0 ? 0.0 : var;
        stats[g * 2] = (float)mean;
        stats[g * 2 + 1] = (float)(1.0 / sqrt(var + (double)eps));
    }
}

// ------------------------------------------------------------------------------------------------ SpatialNorm + SiLU
struct FrameMap {
    int src[32];   // source frame of output frame t (nearest-neighbour temporal index)
};

// out[(t+2), hp, wp, :] = SiLU( ((x - mean_g) * rstd_g * gamma + beta) * Y + B ) on interior positions, 0 on the border ring;
// Y | B = yb[(ts*hl + (h >> lsh))*wl + (w >> lsw), 0:C | C:2C]  (conv_y / conv_b evaluated at latent resolution: a 1x1x1
// convolution commutes with nearest-neighbour upsampling, autoencoder_kl_cogvideox.py:173-187).
__global__ void __launch_bounds__(256, 3) spatialnorm_silu_kernel(const bf16* __restrict__ x, bf16* __restrict__ out,
                                                               const float* __restrict__ stats, const bf16* __restrict__ gamma,
                                                               const bf16* __restrict__ beta, const bf16* __restrict__ yb, long long ldyb,
                                                               FrameMap fm, int T, int H, int W, int C, int G, int hl, int wl, int lsh, int lsw) {
    // A block walks whole padded rows (t, hp); 256 % (C/8) == 0, so a thread always meets the same 8 channels and keeps their
    // GroupNorm scale / offset in registers; no per-element index division; four 16-byte loads in flight per thread.
    const int Hp = H + 2, Wp = W + 2, vpp = C >> 3, cpg = C / G;
    const int v = threadIdx.x % vpp, p0 = threadIdx.x / vpp, pstep = 256 / vpp;
    float ga[8], gc[8];   // nrm = f * ga + gc  with ga = rstd * gamma, gc = beta - mean * rstd * gamma
    {
        float gm[8], bt[8];
        v_unpack8(__ldg(reinterpret_cast<const uint4*>(gamma + v * 8)), gm);
        v_unpack8(__ldg(reinterpret_cast<const uint4*>(beta + v * 8)), bt);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int g = (v * 8 + j) / cpg;
            const float mean = stats[g * 2], rstd = stats[g * 2 + 1];
            ga[j] = rstd * gm[j];
            gc[j] = bt[j] - mean * ga[j];
        }
    }
    const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
    for (int row = blockIdx.x; row < T * Hp; row += gridDim.x) {
        const int t = row / Hp, hp = row - t * Hp;
        const long long base = ((long long)(t + 2) * Hp + hp) * Wp * C + v * 8;
        if (hp == 0 || hp == H + 1) {
            for (int wp = p0; wp < Wp; wp += pstep) *reinterpret_cast<uint4*>(out + base + (long long)wp * C) = zero;
            continue;
        }
        const long long lrow0 = yb ? ((long long)fm.src[t] * hl + ((hp - 1) >> lsh)) * wl : 0;
        if (p0 == 0) {   // the two ring columns of the row
            *reinterpret_cast<uint4*>(out + base) = zero;
            *reinterpret_cast<uint4*>(out + base + (long long)(W + 1) * C) = zero;
        }
        // a thread takes 8 consecutive interior positions at a time: eight 16-byte loads in flight, and the conv_y / conv_b row
        // (shared by 2^lsw neighbours) is fetched only when the latent column changes
        for (int c0 = p0 * 8; c0 < W; c0 += pstep * 8) {
            uint4 u[8];
#pragma unroll
            for (int k = 0; k < 8; ++k)
                u[k] = (c0 + k < W) ? __ldg(reinterpret_cast<const uint4*>(x + base + (long long)(c0 + k + 1) * C)) : zero;
            float yy[8], bb[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                yy[j] = 1.f;   // plain GroupNorm + SiLU when yb == nullptr (the encoder's resnets: spatial_norm_dim=None)
                bb[j] = 0.f;
            }
            int lcol = -1;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int w = c0 + k;   // interior column
                if (w >= W) break;
                if (yb && (w >> lsw) != lcol) {
                    lcol = w >> lsw;
                    const bf16* yr = yb + (lrow0 + lcol) * ldyb + v * 8;
                    v_unpack8(__ldg(reinterpret_cast<const uint4*>(yr)), yy);
                    v_unpack8(__ldg(reinterpret_cast<const uint4*>(yr + C)), bb);
                }
                float f[8];
                v_unpack8(u[k], f);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float nrm = fmaf(f[j], ga[j], gc[j]);
                    const float uu = fmaf(nrm, yy[j], bb[j]);
                    f[j] = __fdividef(uu, 1.0f + __expf(-uu));
                }
                *reinterpret_cast<uint4*>(out + base + (long long)(w + 1) * C) = v_pack8(f);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ nearest upsample (x2 in H, W)
__global__ void __launch_bounds__(256) upsample_nearest_kernel(const bf16* __restrict__ x, bf16* __restrict__ out, FrameMap fm, int T_out,
                                                               int H_in, int W_in, int C) {
    const int H = 2 * H_in, W = 2 * W_in, Hp = H + 2, Wp = W + 2, Hpi = H_in + 2, Wpi = W_in + 2, vpp = C >> 3;
    const long long nvec = (long long)T_out * Hp * Wp * vpp;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
        const int v = (int)(i % vpp);
        long long r = i / vpp;
        const int wp = (int)(r % Wp);
        r /= Wp;
        const int hp = (int)(r % Hp);
        const int t = (int)(r / Hp);
        uint4 o = make_uint4(0u, 0u, 0u, 0u);
        if (hp >= 1 && hp <= H && wp >= 1 && wp <= W)
            o = __ldg(reinterpret_cast<const uint4*>(x + ((long long)((fm.src[t] + 2) * Hpi + ((hp - 1) >> 1) + 1) * Wpi + ((wp - 1) >> 1) + 1) * C +
                                                     v * 8));
        *reinterpret_cast<uint4*>(out + ((long long)((t + 2) * Hp + hp) * Wp + wp) * C + v * 8) = o;
    }
}

// ------------------------------------------------------------------------------------------------ stride-2 subsample
// CogVideoXDownsample3D's conv (D/models/downsampling.py:345-350) is F.pad(0,1,0,1) + a stride-2 3x3 Conv2d: output (y, x) is
// the stride-1 'same' convolution at (2y + 1, 2x + 1) — whose bottom / right neighbours at the edge are the zero ring of the
// padded volume, exactly the reference's padding.  The stride-1 product runs on the implicit-GEMM conv; this kernel keeps the
// odd positions: out[t, y, x, :] = in[t, 2y + 1, 2x + 1, :] (interior coordinates), zero ring.
__global__ void __launch_bounds__(256) subsample2_kernel(const bf16* __restrict__ x, bf16* __restrict__ out, int T, int H_in, int W_in, int C) {
    const int H = H_in / 2, W = W_in / 2, Hp = H + 2, Wp = W + 2, Hpi = H_in + 2, Wpi = W_in + 2, vpp = C >> 3;
    const long long nvec = (long long)T * Hp * Wp * vpp;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
        const int v = (int)(i % vpp);
        long long r = i / vpp;
        const int wp = (int)(r % Wp);
        r /= Wp;
        const int hp = (int)(r % Hp);
        const int t = (int)(r / Hp);
        uint4 o = make_uint4(0u, 0u, 0u, 0u);
        if (hp >= 1 && hp <= H && wp >= 1 && wp <= W)
            o = __ldg(reinterpret_cast<const uint4*>(x + ((long long)((t + 2) * Hpi + 2 * hp) * Wpi + 2 * wp) * C + v * 8));
        *reinterpret_cast<uint4*>(out + ((long long)((t + 2) * Hp + hp) * Wp + wp) * C + v * 8) = o;
    }
}

// ------------------------------------------------------------------------------------------------ conv_out volume -> video
// video[c, f0 + t, h, w] = vol[(t + 2), h + 1, w + 1, c]   (vol has ldc >= Cout channels, video is [Cout, Tv, H, W] bf16)
__global__ void volume_to_video_kernel(const bf16* __restrict__ vol, bf16* __restrict__ video, int T, int H, int W, int ldc, int Cout,
                                       int Tv, int f0) {
    const int Hp = H + 2, Wp = W + 2;
    const long long n = (long long)Cout * T * H * W;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int w = (int)(i % W);
        long long r = i / W;
        const int h = (int)(r % H);
        r /= H;
        const int t = (int)(r % T);
        const int c = (int)(r / T);
        video[(((long long)c * Tv + f0 + t) * H + h) * W + w] = vol[((long long)((t + 2) * Hp + h + 1) * Wp + w + 1) * ldc + c];
    }
}

// ------------------------------------------------------------------------------------------------ video -> uint8 frames
// The host glue after the decoder (D/video_processor.py:89-113 postprocess_video -> D/image_processor.py:227-239 denormalize,
// :196-208 pt_to_numpy, :133-150 numpy_to_pil; D/utils/export_utils.py:177-178 export_to_video) done on the device, with the
// reference's rounding points:  d = bf16(bf16(v / 2) + 0.5)  clamped to [0, 1]  (two bf16 tensor ops),  then in fp32
// d * 255 -> uint8 by truncation (export_to_video's astype) or round-half-even (numpy_to_pil's .round()).
// video [B, 3, F, H, W] bf16 -> frames [B, F, H, W, 3] uint8: 6 bytes read + 3 bytes written per pixel instead of a 4-byte
// fp32 copy per channel to the host.  One thread = 4 consecutive pixels of one frame (8-byte loads per plane, 12-byte store).
__device__ __forceinline__ uint32_t px_u8(float v, int round_mode) {
    float d = bf16_round(bf16_round(v * 0.5f) + 0.5f);
    d = fminf(fmaxf(d, 0.f), 1.f);
    const float y = d * 255.0f;
    return (uint32_t)(round_mode ? rintf(y) : floorf(y));
}
__global__ void __launch_bounds__(256)
video_to_uint8_kernel(const bf16* __restrict__ video, uint8_t* __restrict__ out, int B, int F, long long HW, int round_mode) {
    const long long groups = HW / 4;                       // 4-pixel groups per frame (host checks HW % 4 == 0)
    const long long n = (long long)B * F * groups;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const long long g = i % groups;
        const long long bf = i / groups;
        const int f = (int)(bf % F);
        const int b = (int)(bf / F);
        uint32_t px[12];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const uint2 u = __ldg(reinterpret_cast<const uint2*>(video + (((long long)b * 3 + c) * F + f) * HW) + g);
            px[0 * 3 + c] = px_u8(bf16_lo(u.x), round_mode);
            px[1 * 3 + c] = px_u8(bf16_hi(u.x), round_mode);
            px[2 * 3 + c] = px_u8(bf16_lo(u.y), round_mode);
            px[3 * 3 + c] = px_u8(bf16_hi(u.y), round_mode);
        }
        uint32_t w[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) w[k] = px[4 * k] | (px[4 * k + 1] << 8) | (px[4 * k + 2] << 16) | (px[4 * k + 3] << 24);
        uint32_t* o = reinterpret_cast<uint32_t*>(out + ((long long)(b * F + f) * HW + g * 4) * 3);
        o[0] = w[0]; o[1] = w[1]; o[2] = w[2];
    }
}

// 16 pixels per thread (HW % 16 == 0): two 16-byte loads per channel plane, three 16-byte stores of interleaved RGB — the
// 4-pixel form above moves 8-byte loads and 4-byte stores and reached 0.31 of the HBM copy bandwidth.  Same px_u8 arithmetic.
__global__ void __launch_bounds__(256)
video_to_uint8_x16_kernel(const bf16* __restrict__ video, uint8_t* __restrict__ out, int B, int F, long long HW, int round_mode) {
    const long long groups = HW / 16;
    const long long n = (long long)B * F * groups;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const long long g = i % groups;
        const long long bf = i / groups;
        const int f = (int)(bf % F);
        const int b = (int)(bf / F);
        uint4 u[3][2];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const uint4* src = reinterpret_cast<const uint4*>(video + (((long long)b * 3 + c) * F + f) * HW) + 2 * g;
            u[c][0] = __ldg(src);
            u[c][1] = __ldg(src + 1);
        }
        uint8_t px[48];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const uint32_t w4[4] = {u[c][h].x, u[c][h].y, u[c][h].z, u[c][h].w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    px[(h * 8 + 2 * k) * 3 + c] = (uint8_t)px_u8(bf16_lo(w4[k]), round_mode);
                    px[(h * 8 + 2 * k + 1) * 3 + c] = (uint8_t)px_u8(bf16_hi(w4[k]), round_mode);
                }
            }
        }
        uint4* o = reinterpret_cast<uint4*>(out + ((long long)(b * F + f) * HW + g * 16) * 3);
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            uint4 w;
            w.x = px[16 * q + 0] | (px[16 * q + 1] << 8) | (px[16 * q + 2] << 16) | ((uint32_t)px[16 * q + 3] << 24);
            w.y = px[16 * q + 4] | (px[16 * q + 5] << 8) | (px[16 * q + 6] << 16) | ((uint32_t)px[16 * q + 7] << 24);
            w.z = px[16 * q + 8] | (px[16 * q + 9] << 8) | (px[16 * q + 10] << 16) | ((uint32_t)px[16 * q + 11] << 24);
            w.w = px[16 * q + 12] | (px[16 * q + 13] << 8) | (px[16 * q + 14] << 16) | ((uint32_t)px[16 * q + 15] << 24);
            o[q] = w;
        }
    }
}

// ------------------------------------------------------------------------------------------------ seam blending
// b[o, y, x] = a[o, La - extent + y, x] * (1 - y/extent) + b[o, y, x] * (y/extent) for y < extent along the blended axis,
// with torch's bf16 rounding points (each product and the sum are rounded): autoencoder_kl_cogvideox.py:1284-1298.
// Generic strides (elements): outer index o, blended index y, other index x.
__global__ void blend_kernel(const bf16* __restrict__ a, bf16* __restrict__ b, long long n_outer, int extent, int n_other, int a_len,
                             long long a_so, long long a_sy, long long a_sx, long long b_so, long long b_sy, long long b_sx) {
    const long long n = n_outer * extent * n_other;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int xo = (int)(i % n_other);
        long long r = i / n_other;
        const int y = (int)(r % extent);
        const long long o = r / extent;
        const float wb = (float)((double)y / (double)extent), wa = (float)(1.0 - (double)y / (double)extent);
        const float av = __bfloat162float(a[o * a_so + (long long)(a_len - extent + y) * a_sy + xo * a_sx]);
        bf16* bp = b + o * b_so + (long long)y * b_sy + xo * b_sx;
        const float r1 = bf16_round(av * wa), r2 = bf16_round(__bfloat162float(*bp) * wb);
        *bp = __float2bfloat16_rn(r1 + r2);
    }
}

static inline int grid_for(long long n, int threads) {
    long long b = (n + threads - 1) / threads;
    const long long cap = (long long)sm_count() * 16;
    if (b > cap) b = cap;
    return (int)(b < 1 ? 1 : b);
}

}  // namespace s2v

using namespace s2v;

extern "C" int s2v_vae_latent_rows(const void* z, void* rows, int32_t C, int32_t Tz, int32_t hz, int32_t wz, int32_t f0, int32_t T,
                                   int32_t i0, int32_t j0, int32_t ht, int32_t wt, int32_t ldr, float scale, void* stream) {
    if (!z || !rows) return set_error(S2V_E_BADARG, "s2v_vae_latent_rows: null pointer");
    if (T <= 0 || ht <= 0 || wt <= 0 || f0 < 0 || f0 + T > Tz || i0 < 0 || j0 < 0 || i0 + ht > hz || j0 + wt > wz || ldr < C)
        return set_error(S2V_E_BADARG, "s2v_vae_latent_rows: tile outside the latent");
    int rc = ensure_device();
    if (rc) return rc;
    const long long n = (long long)T * ht * wt * ldr;
    latent_rows_kernel<<<grid_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const bf16*>(z), static_cast<bf16*>(rows), C,
                                                                                       Tz, hz, wz, f0, T, i0, j0, ht, wt, ldr, scale);
    return check_launch("latent_rows_kernel");
}

extern "C" int s2v_vae_latent_im2col(const void* z, void* col, int32_t C, int32_t Tz, int32_t hz, int32_t wz, int32_t f0, int32_t T,
                                     int32_t i0, int32_t j0, int32_t ht, int32_t wt, float scale, void* stream) {
    if (!z || !col) return set_error(S2V_E_BADARG, "s2v_vae_latent_im2col: null pointer");
    if (T <= 0 || ht <= 0 || wt <= 0 || f0 < 0 || f0 + T > Tz || i0 < 0 || j0 < 0 || i0 + ht > hz || j0 + wt > wz)
        return set_error(S2V_E_BADARG, "s2v_vae_latent_im2col: tile outside the latent");
    int rc = ensure_device();
    if (rc) return rc;
    const long long n = (long long)T * (ht + 2) * (wt + 2) * 27 * C;
    latent_im2col_kernel<<<grid_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const bf16*>(z), static_cast<bf16*>(col), C,
                                                                                         Tz, hz, wz, f0, T, i0, j0, ht, wt, scale);
    return check_launch("latent_im2col_kernel");
}

extern "C" int s2v_vae_groupnorm_stats(const void* x, float* partial, float* stats, int32_t T, int32_t H, int32_t W, int32_t C, int32_t G,
                                       int32_t max_blocks, float eps, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (!x || !partial || !stats) return set_error(S2V_E_BADARG, "s2v_vae_groupnorm_stats: null pointer");
    if (T <= 0 || H <= 0 || W <= 0 || C <= 0 || G <= 0 || (C % G) || (C % 8) || (256 % (C / 8)) || max_blocks <= 0)
        return set_error(S2V_E_UNSUPPORTED, "s2v_vae_groupnorm_stats: C must be a multiple of 8 and of G with C/8 dividing 256");
    int rc = ensure_device();
    if (rc) return rc;
    const int lines = T * H;
    int nblk = lines < max_blocks ? lines : max_blocks;
    const int lpb = (lines + nblk - 1) / nblk;
    nblk = (lines + lpb - 1) / lpb;
    gn_partial_kernel<<<nblk, 256, 0, stream>>>(static_cast<const bf16*>(x), partial, T, H, W, C, lpb);
    if ((rc = check_launch("gn_partial_kernel"))) return rc;
    gn_finalize_kernel<<<G, GN_FIN_THREADS, 0, stream>>>(partial, stats, nblk, C, G, (float)((double)T * H * W * (C / G)), eps);
    return check_launch("gn_finalize_kernel");
}

extern "C" int s2v_vae_groupnorm_finalize(const float* stats_partial, float* stats, int32_t row_blocks, int32_t T, int32_t H, int32_t W,
                                          int32_t C, int32_t G, float eps, void* stream) {
    if (!stats_partial || !stats) return set_error(S2V_E_BADARG, "s2v_vae_groupnorm_finalize: null pointer");
    if (row_blocks <= 0 || T <= 0 || H <= 0 || W <= 0 || C <= 0 || G <= 0 || (C % G))
        return set_error(S2V_E_BADARG, "s2v_vae_groupnorm_finalize: bad shape");
    int rc = ensure_device();
    if (rc) return rc;
    gn_finalize_kernel<<<G, GN_FIN_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(stats_partial, stats, row_blocks, C, G, (float)((double)T * H * W * (C / G)), eps);
    return check_launch("gn_finalize_kernel");
}

extern "C" int s2v_vae_spatialnorm_silu(const void* x, void* out, const float* stats, const void* gamma, const void* beta, const void* yb,
                                        int64_t ldyb, const int32_t* frame_src, int32_t T, int32_t H, int32_t W, int32_t C, int32_t G,
                                        int32_t hl, int32_t wl, void* stream) {
    if (!x || !out || !stats || !gamma || !beta || !yb || !frame_src) return set_error(S2V_E_BADARG, "s2v_vae_spatialnorm_silu: null pointer");
    if (T <= 0 || T > 32 || (C % 8) || (C % G) || hl <= 0 || wl <= 0 || (H % hl) || (W % wl))
        return set_error(S2V_E_UNSUPPORTED, "s2v_vae_spatialnorm_silu: T <= 32, C % 8 == 0, H and W multiples of the latent size");
    if (ldyb < 2 * C || (ldyb % 8)) return set_error(S2V_E_BADARG, "s2v_vae_spatialnorm_silu: ldyb must be >= 2*C and a multiple of 8");
    int lsh = 0, lsw = 0;
    while ((hl << lsh) < H) ++lsh;
    while ((wl << lsw) < W) ++lsw;
    if ((hl << lsh) != H || (wl << lsw) != W) return set_error(S2V_E_UNSUPPORTED, "s2v_vae_spatialnorm_silu: scale must be a power of two");
    int rc = ensure_device();
    if (rc) return rc;
    FrameMap fm;
    for (int t = 0; t < 32; ++t) fm.src[t] = t < T ? frame_src[t] : 0;
    if (256 % (C / 8)) return set_error(S2V_E_UNSUPPORTED, "s2v_vae_spatialnorm_silu: C/8 must divide 256");
    const int rows = T * (H + 2);
    spatialnorm_silu_kernel<<<rows < 148 * 8 ? rows : 148 * 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const bf16*>(x), static_cast<bf16*>(out), stats, static_cast<const bf16*>(gamma), static_cast<const bf16*>(beta),
        static_cast<const bf16*>(yb), (long long)ldyb, fm, T, H, W, C, G, hl, wl, lsh, lsw);
    return check_launch("spatialnorm_silu_kernel");
}

extern "C" int s2v_vae_groupnorm_silu(const void* x, void* out, const float* stats, const void* gamma, const void* beta, int32_t T,
                                      int32_t H, int32_t W, int32_t C, int32_t G, void* stream) {
    if (!x || !out || !stats || !gamma || !beta) return set_error(S2V_E_BADARG, "s2v_vae_groupnorm_silu: null pointer");
    if (T <= 0 || T > 32 || (C % 8) || (C % G) || H <= 0 || W <= 0)
        return set_error(S2V_E_UNSUPPORTED, "s2v_vae_groupnorm_silu: T <= 32, C % 8 == 0, C % G == 0");
    int rc = ensure_device();
    if (rc) return rc;
    FrameMap fm;
    for (int t = 0; t < 32; ++t) fm.src[t] = 0;
    if (256 % (C / 8)) return set_error(S2V_E_UNSUPPORTED, "s2v_vae_groupnorm_silu: C/8 must divide 256");
    const int rows = T * (H + 2);
    spatialnorm_silu_kernel<<<rows < 148 * 8 ? rows : 148 * 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const bf16*>(x), static_cast<bf16*>(out), stats, static_cast<const bf16*>(gamma), static_cast<const bf16*>(beta),
        nullptr, 0, fm, T, H, W, C, G, H, W, 0, 0);
    return check_launch("spatialnorm_silu_kernel");
}

extern "C" int s2v_vae_subsample2(const void* x, void* out, int32_t T, int32_t H_in, int32_t W_in, int32_t C, void* stream) {
    if (!x || !out) return set_error(S2V_E_BADARG, "s2v_vae_subsample2: null pointer");
    if (T <= 0 || H_in < 2 || W_in < 2 || (C % 8)) return set_error(S2V_E_UNSUPPORTED, "s2v_vae_subsample2: H, W >= 2 and C % 8 == 0");
    int rc = ensure_device();
    if (rc) return rc;
    const long long nvec = (long long)T * (H_in / 2 + 2) * (W_in / 2 + 2) * (C / 8);
    subsample2_kernel<<<grid_for(nvec, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const bf16*>(x), static_cast<bf16*>(out), T,
                                                                                         H_in, W_in, C);
    return check_launch("subsample2_kernel");
}

extern "C" int s2v_vae_upsample_nearest(const void* x, void* out, const int32_t* frame_src, int32_t T_out, int32_t H_in, int32_t W_in,
                                        int32_t C, void* stream) {
    if (!x || !out || !frame_src) return set_error(S2V_E_BADARG, "s2v_vae_upsample_nearest: null pointer");
    if (T_out <= 0 || T_out > 32 || (C % 8)) return set_error(S2V_E_UNSUPPORTED, "s2v_vae_upsample_nearest: T_out <= 32 and C % 8 == 0");
    int rc = ensure_device();
    if (rc) return rc;
    FrameMap fm;
    for (int t = 0; t < 32; ++t) fm.src[t] = t < T_out ? frame_src[t] : 0;
    const long long nvec = (long long)T_out * (2 * H_in + 2) * (2 * W_in + 2) * (C / 8);
    upsample_nearest_kernel<<<grid_for(nvec, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const bf16*>(x), static_cast<bf16*>(out),
                                                                                               fm, T_out, H_in, W_in, C);
    return check_launch("upsample_nearest_kernel");
}

extern "C" int s2v_vae_volume_to_video(const void* vol, void* video, int32_t T, int32_t H, int32_t W, int32_t ldc, int32_t Cout, int32_t Tv,
                                       int32_t f0, void* stream) {
    if (!vol || !video) return set_error(S2V_E_BADARG, "s2v_vae_volume_to_video: null pointer");
    if (T <= 0 || f0 < 0 || f0 + T > Tv || Cout > ldc) return set_error(S2V_E_BADARG, "s2v_vae_volume_to_video: bad frame range");
    int rc = ensure_device();
    if (rc) return rc;
    const long long n = (long long)Cout * T * H * W;
    volume_to_video_kernel<<<grid_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const bf16*>(vol), static_cast<bf16*>(video),
                                                                                           T, H, W, ldc, Cout, Tv, f0);
    return check_launch("volume_to_video_kernel");
}

extern "C" int s2v_video_to_uint8(const void* video, void* frames, int32_t B, int32_t F, int32_t H, int32_t W, int32_t round_mode,
                                  void* stream) {
    if (!video || !frames) return set_error(S2V_E_BADARG, "s2v_video_to_uint8: null pointer");
    if (B <= 0 || F <= 0 || H <= 0 || W <= 0) return set_error(S2V_E_BADARG, "s2v_video_to_uint8: empty video");
    if (round_mode != 0 && round_mode != 1) return set_error(S2V_E_BADARG, "s2v_video_to_uint8: round_mode is 0 (truncate) or 1 (round half even)");
    const long long HW = (long long)H * W;
    if (HW % 4) return set_error(S2V_E_UNSUPPORTED, "s2v_video_to_uint8: H*W must be a multiple of 4");
    if ((reinterpret_cast<uintptr_t>(video) & 7) || (reinterpret_cast<uintptr_t>(frames) & 3))
        return set_error(S2V_E_BADARG, "s2v_video_to_uint8: video must be 8-byte and frames 4-byte aligned");
    int rc = ensure_device();
    if (rc) return rc;
    if (HW % 16 == 0 && (reinterpret_cast<uintptr_t>(video) % 16 == 0) && (reinterpret_cast<uintptr_t>(frames) % 16 == 0)) {
        const long long n16 = (long long)B * F * (HW / 16);
        video_to_uint8_x16_kernel<<<grid_for(n16, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const bf16*>(video),
                                                                                                    static_cast<uint8_t*>(frames), B, F, HW, round_mode);
        return check_launch("video_to_uint8_x16_kernel");
    }
    const long long n = (long long)B * F * (HW / 4);
    video_to_uint8_kernel<<<grid_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const bf16*>(video),
                                                                                          static_cast<uint8_t*>(frames), B, F, HW, round_mode);
    return check_launch("video_to_uint8_kernel");
}

extern "C" int s2v_vae_blend(const void* a, void* b, int64_t n_outer, int32_t extent, int32_t n_other, int32_t a_len, int64_t a_so,
                             int64_t a_sy, int64_t a_sx, int64_t b_so, int64_t b_sy, int64_t b_sx, void* stream) {
    if (!a || !b) return set_error(S2V_E_BADARG, "s2v_vae_blend: null pointer");
    if (extent <= 0 || extent > a_len || n_outer <= 0 || n_other <= 0) return set_error(S2V_E_BADARG, "s2v_vae_blend: bad extent");
    int rc = ensure_device();
    if (rc) return rc;
    const long long n = n_outer * extent * n_other;
    blend_kernel<<<grid_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const bf16*>(a), static_cast<bf16*>(b), n_outer, extent,
                                                                                 n_other, a_len, a_so, a_sy, a_sx, b_so, b_sy, b_sx);
    return check_launch("blend_kernel");
}
