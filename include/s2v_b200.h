/*
 * s2v_b200 — C ABI of the B200-native (sm_100a) kernels behind the CogVideoX subject-to-video denoising loop.
 *
 * The reference (carpedkm/disentangled-subject-to-vid) is pure Python on a diffusers fork and has NO native FFI;
 * each entry point below replaces the group of PyTorch library calls named in its comment (file:line under
 * /root/reference, S/ = src/, D/ = diffusers/src/diffusers/).  INTEGRATION.md shows the ctypes binding a
 * maintainer adds on the reference side.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host; bf16 = 16-bit brain float; row-major;
 *     `ld*` are leading dimensions in ELEMENTS;
 *   - the caller owns every buffer (inputs, outputs, workspaces); the library allocates nothing and keeps no
 *     tensor state; all launches go to `stream` (a cudaStream_t passed as void*), are asynchronous, do not
 *     synchronise the host and are CUDA-graph-capture safe;
 *   - return value: 0 = ok, negative = bad argument / unsupported shape (S2V_E_*), positive = cudaError_t.
 *     Nothing throws or exits across the ABI; s2v_last_error() returns a thread-local description.
 *   - there is no CPU fallback and no other backend: on a machine without an sm_100 device every compute
 *     entry point returns S2V_E_NO_DEVICE.
 *
 * Token layout: every activation of the transformer is ONE buffer [B, S, D] with rows ordered
 * [text (L) | reference image (n_ref) | video (F*n)] — the order of the reference's joint attention
 * (D/models/attention_processor.py:2039), so its torch.cat / split copies disappear.
 */
#ifndef S2V_B200_H
#define S2V_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define S2V_ABI_VERSION 1

#if defined(__GNUC__)
#define S2V_API __attribute__((visibility("default")))
#else
#define S2V_API
#endif

#define S2V_E_BADARG (-1)
#define S2V_E_UNSUPPORTED (-2)
#define S2V_E_NO_DEVICE (-3)
#define S2V_E_DRIVER (-4)

/* epilogues of s2v_linear */
#define S2V_EPI_BIAS 0          /* out = bf16(alpha*acc + bias)                                   */
#define S2V_EPI_BIAS_GELU 1     /* out = bf16(gelu_tanh(acc + bias))                              */
#define S2V_EPI_GATE_RESIDUAL 2 /* out = bf16(out + gate[batch(row), seg(row), col]*(acc + bias)) */

S2V_API int s2v_abi_version(void);
S2V_API const char* s2v_last_error(void);
/* 0 if device `dev` is sm_100 (B200); S2V_E_NO_DEVICE otherwise. */
S2V_API int s2v_device_check(int dev);
/* Bytes of caller-owned activation workspace one transformer forward needs (the library allocates nothing):
 * h, xn, att [B,S,D] + qkv [B,S,3D] + ffh [B,S,ff_dim] in bf16, the LoRA down-projection scratch [B*S, max(lora_cols, 8)] bf16 and
 * the modulation table [n_mod, B, 6D] fp32, each rounded up to 256 bytes.  Host-only arithmetic (no device needed); negative
 * on bad arguments.  The buffers may be laid out in one allocation in that order. */
S2V_API int64_t s2v_workspace_bytes(int32_t B, int32_t S, int32_t D, int32_t ff_dim, int32_t lora_cols, int32_t n_mod);

/* ------------------------------------------------------------------------------------------------ linear + LoRA
 * y[M,N] = x[M,K] * w[N,K]^T (+ bias) (+ lora_t[M, g*r : (g+1)*r] * lora_b[N, r]^T), tcgen05 + TMA, fp32 accumulate,
 * with a fused epilogue.  `lora_t` is the already down-projected and scaled activation  s * x * A^T  (a first
 * s2v_linear call with alpha = s and w = A); the up-projection B rides along as `lora_r` extra K columns of the main
 * GEMM, so the LoRA update costs no extra pass over y.  g = (column / lora_group_n) lets q/k/v share one call.
 * Replaces: nn.Linear + peft lora.Linear at D/models/attention_processor.py:2049-2051,2090 (to_q/to_k/to_v/to_out),
 * D/models/attention.py:1241-1242 (ff.net.0.proj, ff.net.2), D/models/embeddings.py:411,416 (text_proj, proj),
 * cogvideox_transformer_3d.py:543 (proj_out); GELU-tanh of D/models/activations.py:88-89; the gated residual adds of
 * cogvideox_transformer_3d.py:165-167,182-184.
 */
typedef struct {
    const void* x;      int64_t ldx;   /* [M,K] bf16 */
    const void* w;      int64_t ldw;   /* [N,K] bf16 */
    const void* bias;                  /* [N] bf16 or NULL */
    const void* lora_t; int64_t ldt;   /* [M, groups*lora_r] bf16 or NULL */
    const void* lora_b; int64_t ldb;   /* [N, lora_r] bf16 */
    int32_t lora_r;                    /* 0 = no LoRA; multiple of 8 */
    int32_t lora_group_n;              /* columns per LoRA group (N if a single group) */
    void* out;          int64_t ldo;   /* [M,N] bf16 (read-modify-write for GATE_RESIDUAL) */
    int32_t M, N, K;                   /* K multiple of 8 */
    int32_t epilogue;                  /* S2V_EPI_* */
    float alpha;
    /* GATE_RESIDUAL only: modulation table [B, mod_stride] fp32; gate vector for a row is
       mod + batch*mod_stride + (seq < text_len ? gate_off_text : gate_off_other), batch = row / rows_per_batch */
    const float* mod;   int32_t mod_stride, gate_off_text, gate_off_other, rows_per_batch, text_len;
} s2v_linear_args;
S2V_API int s2v_linear(const s2v_linear_args* a, void* stream);

/* ------------------------------------------------------------------------------------------------ joint attention
 * o[b, s, h*64:(h+1)*64] = softmax(q k^T / 8) v over ALL S rows (text|ref|video), no mask, head_dim 64.
 * qkv is the fused projection output [B, S, 3*H*64] bf16 (q | k | v along the last dim, heads contiguous);
 * o is [B, S, H*64] bf16 — directly the A operand of the out-projection.  tcgen05 flash attention: S and P live in
 * TMEM, K/V tiles are staged by TMA, fp32 online softmax.
 * Replaces F.scaled_dot_product_attention + the two transposes at D/models/attention_processor.py:2056-2058,2083-2088.
 */
S2V_API int s2v_attn_fwd(const void* qkv, void* o, int32_t B, int32_t S, int32_t H, float softmax_scale, void* stream);

/* ------------------------------------------------------------------------------------------------ AdaLN-Zero
 * out[b,s,:] = LayerNorm_D(x[b,s,:]; w, bias, eps) * (1 + scale) + shift, where (shift, scale) come from the
 * modulation table mod[B, mod_stride] fp32 at offsets (shift_off_text, scale_off_text) for s < text_len and
 * (shift_off_other, scale_off_other) otherwise.  CogVideoXLayerNormZero (D/models/normalization.py:479-482): video and
 * reference rows share chunks 0/1, text rows use chunks 3/4 (the `enable_lora` context there is a no-op).
 */
S2V_API int s2v_adaln_modulate(const void* x, void* out, const void* ln_w, const void* ln_b, const float* mod,
                       int32_t mod_stride, int32_t shift_off_text, int32_t scale_off_text, int32_t shift_off_other,
                       int32_t scale_off_other, int32_t B, int32_t S, int32_t D, int32_t text_len, float eps,
                       void* stream);

/* Final norms: out = LN2(LN1(x)) * (1 + scale) + shift on rows [row0, S) of every batch, written densely as
 * [B, S-row0, D].  nn.LayerNorm norm_final + AdaLayerNorm norm_out (cogvideox_transformer_3d.py:536-542,
 * D/models/normalization.py:70-81: chunk order shift, scale). */
S2V_API int s2v_final_norm(const void* x, void* out, const void* ln1_w, const void* ln1_b, const void* ln2_w,
                   const void* ln2_b, const float* mod, int32_t mod_stride, int32_t shift_off, int32_t scale_off,
                   int32_t B, int32_t S, int32_t row0, int32_t D, float eps, void* stream);

/* ------------------------------------------------------------------------------------------------ q/k LayerNorm + RoPE
 * In place on the q and k thirds of qkv [B, S, 3*H*64]: per-head LayerNorm(64, eps, affine shared across heads), then
 * for rows s >= text_len the interleaved-pair rotation out = x*cos + rot(x)*sin with table row (s - text_len) of
 * cos/sin [S - text_len, 64] fp32 (reference-image rows first, then video rows — the order of the pipeline's table,
 * S/custom_cogvideox_pipe.py:228-235).  cos == NULL skips RoPE (CogVideoX-2B).
 * Replaces attn.norm_q / norm_k and apply_rotary_emb (D/models/attention_processor.py:2060-2080,
 * D/models/embeddings.py:759-778). */
S2V_API int s2v_qk_norm_rope(void* qkv, const void* nq_w, const void* nq_b, const void* nk_w, const void* nk_b,
                     const float* cos, const float* sin, int32_t B, int32_t S, int32_t H, int32_t text_len, float eps,
                     void* stream);

/* ------------------------------------------------------------------------------------------------ small-batch linear
 * out[b,n] = beta*out[b,n] + alpha*(sum_k w[n,k]*f(x[b,k]) + bias[n]),  f = identity (act_in=0) or SiLU (act_in=1);
 * x/out fp32, w/bias bf16, B <= 8.  Used for the timestep MLP (D/models/embeddings.py:860-875), the 6*D AdaLN
 * modulation vectors (D/models/normalization.py:473-476) incl. their LoRA pair, and norm_out.linear. */
S2V_API int s2v_small_linear(const float* x, int64_t ldx, const void* w, int64_t ldw, const void* bias, float* out,
                     int64_t ldo, int32_t B, int32_t N, int32_t K, int32_t act_in, float alpha, float beta,
                     int32_t round_bf16, void* stream);

/* The same function for `count` independent problems in ONE launch (descriptors in DEVICE memory, built once per weight packing):
 * all AdaLN modulation linears of a step (D/models/normalization.py:473-476 for every block, norm_out.linear) and, in a second
 * launch, their LoRA up-projections — 2 launches instead of 3 per linear.  A descriptor with x == NULL reads x_default / ldx_default
 * (the time embedding of this step).  max_n = the largest N of the batch.  Per problem the arithmetic is s2v_small_linear's. */
typedef struct {
    const void* w;        /* [N, K] bf16 */
    int64_t ldw;
    const void* bias;     /* [N] bf16 or NULL */
    float* out;           /* [B, N] fp32 */
    int64_t ldo;
    const float* x;       /* [B, K] fp32, or NULL = the launch's x_default */
    int64_t ldx;
    int32_t N, K, act_in, round_bf16;
    float alpha, beta;
} s2v_small_linear_desc;
S2V_API int s2v_small_linear_batch(const s2v_small_linear_desc* descs, int32_t count, int32_t max_n, const float* x_default,
                                   int64_t ldx_default, int32_t B, void* stream);

/* Sinusoidal timestep embedding, flip_sin_to_cos=True (D/models/embeddings.py:27-78): out[b, 0:D/2] = cos(t_b * f_i),
 * out[b, D/2:] = sin(t_b * f_i); freqs[i] = exp(-ln(10000) * i / (D/2 - freq_shift)) is tabulated by the host in fp32
 * with the reference's own expression so the argument is bit-identical; rounded through bf16 when round_bf16
 * (cogvideox_transformer_3d.py:490). */
S2V_API int s2v_timestep_sinusoid(const float* t, const float* freqs, float* out, int32_t B, int32_t D, int32_t round_bf16,
                                  void* stream);

/* ------------------------------------------------------------------------------------------------ patchify / unpatchify
 * patchify: latents [NB, C, H, W] bf16 -> rows [NB*(H/p)*(W/p), C*p*p] bf16 (k = c*p*p + dy*p + dx), the im2col of
 * the k=p, stride=p Conv2d patch embedding (D/models/embeddings.py:414-419); the conv itself is an s2v_linear call.
 * unpatchify: tokens [NB, (H/p)*(W/p), C*p*p] bf16 -> [NB, C, H, W] bf16 (cogvideox_transformer_3d.py:549-551). */
S2V_API int s2v_patchify(const void* latents, void* rows, int32_t NB, int32_t C, int32_t H, int32_t W, int32_t p, void* stream);
S2V_API int s2v_unpatchify(const void* tokens, void* latents, int32_t NB, int32_t C, int32_t H, int32_t W, int32_t p,
                   void* stream);

/* dst[b, row0 + r, :] += table[r, :] for r < R (bf16; CogVideoX-2B sincos positional embedding on the video rows,
 * D/models/embeddings.py:433-446). */
S2V_API int s2v_add_rows(void* dst, const void* table, int32_t B, int32_t S, int32_t D, int32_t row0, int32_t R, void* stream);

/* ------------------------------------------------------------------------------------------------ CFG + DDIM step
 * latents' = bf16( a*x + b*( sa*x - sb*v ) ),  v = u + g*(t - u) in fp32 from noise_pred [2P, n] bf16 (uncond first),
 * with the reference's rounding points: sa*x and a*x are rounded to bf16 before they meet fp32 terms, every fp32
 * op is individually rounded (no FMA contraction) — bit-exact with S/custom_cogvideox_pipe.py:266-296 +
 * D/schedulers/scheduling_ddim_cogvideox.py:383-394.  The four coefficients are computed by the host in fp64 exactly
 * as the reference and passed as fp32.  x0_out (fp32, optional) receives pred_original_sample. */
S2V_API int s2v_cfg_ddim_step(const void* noise_pred, const void* latents, void* latents_out, float* x0_out, int64_t n_per_half,
                      float guidance, float sqrt_alpha, float sqrt_beta, float a_coef, float b_coef, void* stream);

/* Plain DDIM v-prediction step on an fp32 model output and a bf16 sample (the scheduler.step() surface,
 * D/schedulers/scheduling_ddim_cogvideox.py:365-394): prev_out, x0_out fp32, same rounding points as above. */
S2V_API int s2v_ddim_step(const float* model_output, const void* sample, float* prev_out, float* x0_out, int64_t n,
                          float sqrt_alpha, float sqrt_beta, float a_coef, float b_coef, void* stream);

/* DPM-Solver++ SDE multistep update (the `CogVideoXDPMScheduler` branch of the loop, S/custom_cogvideox_pipe.py:288-295;
 * D/schedulers/scheduling_dpm_cogvideox.py:383-439), bit-exact with torch on CUDA:
 *   x0 = bf16(sa*x) - sb*v;  d = old_x0 ? m2*x0 - m3*old_x0 : x0;  prev = (bf16(m0*x) - m1*d) + bf16(m_noise*noise)
 * cfg_input = 1: model_output is the bf16 [2, n] CFG pair (uncond first), v = u + g*(t-u), prev_out is bf16 (the pipe's cast);
 * cfg_input = 0: model_output is fp32 [n], prev_out fp32 (scheduler.step surface).  noise is the bf16 randn draw the caller
 * made with the reference's generator protocol (two draws per second-order step; the second is the one used).  The host
 * computes the seven coefficients in fp64 exactly as get_variables / get_mult (:306-328).  x0_out fp32 is mandatory (it is
 * the next step's old_pred_original_sample). */
S2V_API int s2v_dpm_step(const void* model_output, int32_t cfg_input, const void* sample, const float* old_x0, const void* noise,
                         void* prev_out, float* x0_out, int64_t n, float guidance, float sqrt_alpha, float sqrt_beta, float m0, float m1,
                         float m2, float m3, float m_noise, void* stream);

/* ------------------------------------------------------------------------------------------------ VAE decoder (row V)
 * Activations of the 3D causal VAE decoder live in padded channels-last VOLUMES  [t_pad + T, Hp, Wp, C]  bf16 with
 * Hp = H + 2, Wp = W + 2 (one zero pixel all round) and t_pad = 2 leading frames that hold the temporal context of a
 * causal convolution (the reference's conv_cache, or copies of the first frame).  Row index r = (t*Hp + hp)*Wp + wp.
 *
 * s2v_conv_gemm: implicit-GEMM convolution on tcgen05 — no im2col is materialised; each of the `taps` kernel taps is a
 * block of K columns whose A rows are the SAME volume shifted by a constant row offset (TMA coordinates), so a
 * 3x3x3 causal conv (taps 27), a per-frame 3x3 conv (taps 9) and a 1x1x1 conv (taps 1) are one kernel:
 *   out[t_pad*Hp*Wp + m, :] = bias + res[...] + sum_tap x[t_pad*Hp*Wp + m + off(tap), :] * w[:, tap*cin : (tap+1)*cin]^T
 * for m in [0, T*Hp*Wp); border positions (hp or wp on the zero ring) are stored as 0 so the result is again a valid
 * padded volume.  w is [cout, taps*cin] with tap = (dt*3 + dh)*3 + dw (torch weight .permute(0,2,3,4,1)).
 * Replaces CogVideoXCausalConv3d / CogVideoXSafeConv3d (autoencoder_kl_cogvideox.py:43-66,129-137), the Conv2d of
 * CogVideoXUpsample3D (upsampling.py:407-410) and the residual add of CogVideoXResnetBlock3D (:318). */
typedef struct {
    const void* x;   int64_t ldx;    /* input volume, [t_pad + T, Hp, Wp, cin] bf16 */
    const void* w;   int64_t ldw;    /* [cout, taps*cin] bf16 */
    const void* bias;                /* [cout] bf16 or NULL */
    const void* res; int64_t ldres;  /* optional residual volume (same geometry as out) or NULL */
    void* out;       int64_t ldo;    /* output volume, [t_pad + T, Hp, Wp, cout] bf16 */
    int32_t T, t_pad, Hp, Wp, cin, cout, taps;
} s2v_conv_args;
S2V_API int s2v_conv_gemm(const s2v_conv_args* c, void* stream);
/* The same convolution with the GroupNorm statistics of its OUTPUT fused into the epilogue: per row-block partial sums
 * stats_partial[row_blocks, cout, 2] fp32 (sum, sum of squares of the bf16 values that are stored, over the T real frames; the zero
 * border ring adds nothing) in a fixed order (deterministic); *row_blocks (host) receives the number of row blocks written
 * (<= ceil(T*Hp*Wp / 128), which sizes the caller's buffer).  s2v_vae_groupnorm_finalize turns them into (mean, rstd) per group.
 * Replaces the separate read of the volume by s2v_vae_groupnorm_stats when the normalised tensor is a convolution output — every
 * nn.GroupNorm input of CogVideoXDecoder3D is (autoencoder_kl_cogvideox.py:297,306,974). */
S2V_API int s2v_conv_gemm_stats(const s2v_conv_args* c, float* stats_partial, int32_t* row_blocks, void* stream);
S2V_API int s2v_vae_groupnorm_finalize(const float* stats_partial, float* stats, int32_t row_blocks, int32_t T, int32_t H, int32_t W,
                                       int32_t C, int32_t G, float eps, void* stream);

/* Latent tile -> GEMM operands.  z is ONE sample [C, Tz, hz, wz] bf16; the tile is frames [f0, f0+T), rows [i0, i0+ht),
 * cols [j0, j0+wt); values are multiplied by `scale` (= 1/scaling_factor, pipeline_cogvideox.py:348) in fp32 and rounded.
 * latent_rows: channels-last rows [T*ht*wt, ldr] (columns >= C zero) — the A operand of the fused conv_y|conv_b 1x1x1 GEMM
 * of every CogVideoXSpatialNorm3D (autoencoder_kl_cogvideox.py:183-184).
 * latent_im2col: [T*(ht+2)*(wt+2), 27*C] — conv_in's A operand over the padded output grid; frame index max(f0+t+dt-2, 0)
 * reproduces CogVideoXCausalConv3d's cache / first-frame replication (:120-127). */
S2V_API int s2v_vae_latent_rows(const void* z, void* rows, int32_t C, int32_t Tz, int32_t hz, int32_t wz, int32_t f0, int32_t T,
                                int32_t i0, int32_t j0, int32_t ht, int32_t wt, int32_t ldr, float scale, void* stream);
S2V_API int s2v_vae_latent_im2col(const void* z, void* col, int32_t C, int32_t Tz, int32_t hz, int32_t wz, int32_t f0, int32_t T,
                                  int32_t i0, int32_t j0, int32_t ht, int32_t wt, float scale, void* stream);

/* GroupNorm(G groups, eps) statistics of the T real frames (interior positions) of a volume: stats[g] = (mean, rstd) fp32.
 * `partial` is caller-owned scratch of max_blocks*C*2 floats; the two-stage reduction has a fixed order (deterministic).
 * nn.GroupNorm inside CogVideoXSpatialNorm3D (autoencoder_kl_cogvideox.py:163,186). */
S2V_API int s2v_vae_groupnorm_stats(const void* x, float* partial, float* stats, int32_t T, int32_t H, int32_t W, int32_t C, int32_t G,
                                    int32_t max_blocks, float eps, void* stream);

/* out = SiLU( GroupNorm(x) * conv_y(zq') + conv_b(zq') ) written as a padded volume (border ring zero).  yb is the fused
 * conv_y|conv_b output at LATENT resolution: rows of [2C] values with row stride ldyb elements (ldyb > 2C when the tables of
 * all norm layers of a decoder call come out of ONE GEMM and a layer reads its column slice); frame_src[t] (host array, T <= 32) is the latent frame that
 * nearest-neighbour interpolation assigns to output frame t (first frame kept separate when T is odd and > 1,
 * autoencoder_kl_cogvideox.py:173-181); H/hl and W/wl must be powers of two.
 * Replaces CogVideoXSpatialNorm3D.forward + the SiLU that follows it (:167-188, :297, :306, :974). */
S2V_API int s2v_vae_spatialnorm_silu(const void* x, void* out, const float* stats, const void* gamma, const void* beta, const void* yb,
                                     int64_t ldyb, const int32_t* frame_src, int32_t T, int32_t H, int32_t W, int32_t C, int32_t G,
                                     int32_t hl, int32_t wl, void* stream);

/* Nearest 2x upsampling in H and W with the temporal index map frame_src[t] (host array, T_out <= 32), written as the padded
 * input volume of the per-frame 3x3 conv.  The interpolate half of CogVideoXUpsample3D.forward (upsampling.py:385-405). */
S2V_API int s2v_vae_upsample_nearest(const void* x, void* out, const int32_t* frame_src, int32_t T_out, int32_t H_in, int32_t W_in,
                                     int32_t C, void* stream);

/* video[c, f0 + t, h, w] = vol[2 + t, h + 1, w + 1, c] for c < Cout: conv_out's volume (ldc channels) -> [Cout, Tv, H, W] bf16. */
S2V_API int s2v_vae_volume_to_video(const void* vol, void* video, int32_t T, int32_t H, int32_t W, int32_t ldc, int32_t Cout, int32_t Tv,
                                    int32_t f0, void* stream);

/* In-place seam ramp of tiled_decode (blend_v / blend_h, autoencoder_kl_cogvideox.py:1284-1298) on bf16 tensors with
 * arbitrary element strides (outer, blended axis, other axis): b[o,y,x] = a[o, a_len-extent+y, x]*(1-y/extent) + b[o,y,x]*(y/extent),
 * with torch's bf16 rounding points. */
/* Encoder side (reference-image VAE encode).  Plain GroupNorm(G) + SiLU of the encoder's resnets
 * (D/models/autoencoders/autoencoder_kl_cogvideox.py:241-243, 292-305 with zq = None) on a padded channels-last volume, statistics
 * from s2v_vae_groupnorm_stats; and the odd-position subsample that turns the stride-1 3x3 convolution into
 * CogVideoXDownsample3D's F.pad(0,1,0,1) + stride-2 Conv2d (D/models/downsampling.py:345-350): out is [T+2, H_in/2+2, W_in/2+2, C]. */
S2V_API int s2v_vae_groupnorm_silu(const void* x, void* out, const float* stats, const void* gamma, const void* beta, int32_t T,
                                   int32_t H, int32_t W, int32_t C, int32_t G, void* stream);
S2V_API int s2v_vae_subsample2(const void* x, void* out, int32_t T, int32_t H_in, int32_t W_in, int32_t C, void* stream);
/* Device-side post-processing of the decoded video (replaces the host chain D/video_processor.py:89-113 postprocess_video ->
 * D/image_processor.py:227-239 denormalize -> :196-208 pt_to_numpy -> D/utils/export_utils.py:177-178 `(frame * 255).astype(uint8)`,
 * or :133-150 numpy_to_pil's `(images * 255).round().astype("uint8")`), bit-exact with the reference's bf16 / fp32 rounding points.
 * video [B, 3, F, H, W] bf16 in [-1, 1] -> frames [B, F, H, W, 3] uint8.  round_mode 0 = truncate (export_to_video),
 * 1 = round half to even (PIL path).  H*W must be a multiple of 4. */
S2V_API int s2v_video_to_uint8(const void* video, void* frames, int32_t B, int32_t F, int32_t H, int32_t W, int32_t round_mode,
                               void* stream);
S2V_API int s2v_vae_blend(const void* a, void* b, int64_t n_outer, int32_t extent, int32_t n_other, int32_t a_len, int64_t a_so,
                          int64_t a_sy, int64_t a_sx, int64_t b_so, int64_t b_sy, int64_t b_sx, void* stream);

/* ------------------------------------------------------------------------------------------------ stage wrappers
 * Thin named entry points (the per-stage ABI proposed in SURVEY.md §8b); each forwards to s2v_linear. */
S2V_API int s2v_qkv_lora(const s2v_linear_args* a, void* stream);                     /* K1: to_q|to_k|to_v + LoRA, bias      */
/* K1 + K2 + K3 in one launch: the fused q|k|v projection (bias, LoRA) whose epilogue applies the per-head LayerNorm(64) and
 * the interleaved RoPE of s2v_qk_norm_rope to the q and k head vectors before they are stored (bit-identical to running
 * s2v_qkv_lora followed by s2v_qk_norm_rope; saves one read + write of 2/3 of qkv per block). */
typedef struct {
    const void *nq_w, *nq_b, *nk_w, *nk_b;   /* [64] bf16 each */
    const float *cos, *sin;                  /* [S - text_len, 64] fp32 or both NULL */
    int32_t S, H, text_len;
    float eps;
} s2v_qk_norm_args;
S2V_API int s2v_qkv_lora_norm_rope(const s2v_linear_args* a, const s2v_qk_norm_args* qk, void* stream);
S2V_API int s2v_outproj_lora_gate_residual(const s2v_linear_args* a, void* stream);   /* K5+K8                                */
S2V_API int s2v_ffn_up_gelu_lora(const s2v_linear_args* a, void* stream);             /* K10 first half                       */
S2V_API int s2v_ffn_down_lora_gate_residual(const s2v_linear_args* a, void* stream);  /* K10 second half + K8                 */

/* ------------------------------------------------------------------------------------------------ T5 prompt encoder
 * The pieces of transformers' modeling_t5.py around the projections (which go through s2v_linear), for the encoder call of
 * D/pipelines/cogvideo/pipeline_cogvideox.py:197-237 (S/inference.py:185-189 loads T5EncoderModel): 226 tokens, no attention mask.
 *   s2v_gather_rows   out[i,:] = table[ids[i],:]                          nn.Embedding (shared.weight); ids int64 on the device
 *   s2v_rmsnorm       out = w * bf16(x * rsqrt(mean(x^2) + eps))          T5LayerNorm.forward (variance in fp32, both roundings)
 *   s2v_gated_gelu    out[m,f] = bf16(gelu_new(g[m,f])) * g[m,F+f]        T5DenseGatedActDense with wi_0 | wi_1 stacked along N
 *   s2v_t5_attention  softmax(bf16(q k^T) + bias) v per (batch, head), head_dim 64, NO 1/sqrt(d), bias [H,S,S] bf16 (the relative
 *                     position table gathered by bucket), qkv [B,S,3*H*64] (q|k|v), out [B,S,H*64]; S <= 512     T5Attention.forward */
S2V_API int s2v_gather_rows(const void* table, const int64_t* ids, void* out, int32_t n, int32_t D, int32_t V, void* stream);
S2V_API int s2v_rmsnorm(const void* x, const void* w, void* out, int32_t rows, int32_t D, float eps, void* stream);
S2V_API int s2v_gated_gelu(const void* g, void* out, int64_t M, int32_t F, void* stream);
S2V_API int s2v_t5_attention(const void* qkv, const void* bias, void* out, int32_t B, int32_t S, int32_t H, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* S2V_B200_H */
