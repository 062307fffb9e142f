"""CPU oracle for the CogVideoX subject-to-video denoising hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under `oracle/` is part of the product: only
`tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference`
legs of `bench.py` may import it, and there only as the checker / CPU timing
arm.  The product path (`disentangled-subject-to-vid_b200`) never imports this module
and fails loudly when its CUDA library is missing.

What this is: a functional restatement, in plain torch-CPU ops (fp32 by default),
of the reference algorithm (carpedkm/disentangled-subject-to-vid @ 4901847a).  Every
function cites the reference file:line it follows.  Shorthand:

    S/  = /root/reference/src/
    D/  = /root/reference/diffusers/src/diffusers/

Parity pinning: the reference ships NO tests and NO golden vectors (its
.gitignore strips `test*`), so the oracle is pinned against outputs of the
reference's own modules executed in the build container; the fixtures and the
script that made them are committed under tests/golden/ (make_golden.py).
The LoRA arithmetic lives in un-vendored, un-pinned `peft` (installer.sh:5,
diffusers requires peft>=0.6.0 — D/utils/constants.py:24); it is restated here
from PEFT's published `lora.Linear` / `lora.Conv2d` forward
(`y = base(x) + scaling * B(A(dropout(x)))`, scaling = lora_alpha / r) and is
"parity unpinned" by the reference itself.

Parameters are passed as a flat dict keyed exactly like the reference
`state_dict()` (e.g. "transformer_blocks.0.attn1.to_q.weight"), with LoRA
factors under "<module>.lora_A.weight" / "<module>.lora_B.weight" — the key
schema of `pytorch_lora_weights_transformer.safetensors` after the
"transformer." prefix is stripped (S/inference.py:68-105).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Params = Dict[str, torch.Tensor]


# --------------------------------------------------------------------------- config
@dataclass
class TransformerConfig:
    """Mirror of the `register_to_config` fields used on the hot path
    (D/models/transformers/cogvideox_transformer_3d.py:252-280)."""

    num_attention_heads: int = 30
    attention_head_dim: int = 64
    in_channels: int = 16
    out_channels: int = 16
    time_embed_dim: int = 512
    text_embed_dim: int = 4096
    num_layers: int = 30
    patch_size: int = 2
    temporal_compression_ratio: int = 4
    max_text_seq_length: int = 226
    norm_eps: float = 1e-5
    spatial_interpolation_scale: float = 1.875
    temporal_interpolation_scale: float = 1.0
    use_rotary_positional_embeddings: bool = False
    flip_sin_to_cos: bool = True
    freq_shift: int = 0
    lora_rank: int = 0
    lora_alpha: float = 0.0

    @property
    def inner_dim(self) -> int:
        return self.num_attention_heads * self.attention_head_dim

    @property
    def lora_scale(self) -> float:
        return (self.lora_alpha / self.lora_rank) if self.lora_rank else 0.0


def config_2b(**kw) -> TransformerConfig:
    # DX/scripts/convert_cogvideox_to_diffusers.py:205-216 (2B: 30 heads, 30 layers, sincos pos-emb)
    return TransformerConfig(num_attention_heads=30, num_layers=30, use_rotary_positional_embeddings=False, **kw)


def config_5b(**kw) -> TransformerConfig:
    # same file: 5B = 48 heads, 42 layers, rotary
    return TransformerConfig(num_attention_heads=48, num_layers=42, use_rotary_positional_embeddings=True, **kw)


# --------------------------------------------------------------------------- LoRA (peft restatement)
def lora_linear(p: Params, name: str, x: torch.Tensor, scale: float) -> torch.Tensor:
    """nn.Linear, optionally wrapped by a PEFT LoRA layer (S/inference.py:218-225 injects
    r=128, alpha=64 => scaling 0.5 on to_q/to_k/to_v/to_out.0/proj/text_proj/norm{1,2}.linear/ff.net.2).
    PEFT lora.Linear.forward: result = base(x); result += lora_B(lora_A(dropout(x))) * scaling."""
    y = F.linear(x, p[name + ".weight"], p.get(name + ".bias"))
    a = p.get(name + ".lora_A.weight")
    if a is not None:
        b = p[name + ".lora_B.weight"]
        y = y + F.linear(F.linear(x, a), b) * scale
    return y


def lora_conv2d(p: Params, name: str, x: torch.Tensor, stride: int, scale: float) -> torch.Tensor:
    """nn.Conv2d wrapped by PEFT lora.Conv2d: A = Conv2d(in, r, kernel, stride), B = Conv2d(r, out, 1x1)."""
    y = F.conv2d(x, p[name + ".weight"], p.get(name + ".bias"), stride=stride)
    a = p.get(name + ".lora_A.weight")
    if a is not None:
        b = p[name + ".lora_B.weight"]
        y = y + F.conv2d(F.conv2d(x, a, None, stride=stride), b) * scale
    return y


# --------------------------------------------------------------------------- embeddings
def timestep_sinusoid(timesteps: torch.Tensor, dim: int, flip_sin_to_cos: bool = True, freq_shift: float = 0.0):
    """D/models/embeddings.py:27-78 (get_timestep_embedding), always fp32."""
    half = dim // 2
    exponent = -math.log(10000) * torch.arange(0, half, dtype=torch.float32)
    exponent = exponent / (half - freq_shift)
    emb = timesteps[:, None].float() * torch.exp(exponent)[None, :]
    emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=-1)
    if flip_sin_to_cos:
        emb = torch.cat([emb[:, half:], emb[:, :half]], dim=-1)
    return emb


def time_embedding(p: Params, cfg: TransformerConfig, timestep: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    """D/models/transformers/cogvideox_transformer_3d.py:484-491 + D/models/embeddings.py:831-876.
    time_embedding.* is NOT a LoRA target."""
    t_emb = timestep_sinusoid(timestep, cfg.inner_dim, cfg.flip_sin_to_cos, cfg.freq_shift).to(dtype)
    h = F.linear(t_emb, p["time_embedding.linear_1.weight"], p["time_embedding.linear_1.bias"])
    h = F.silu(h)
    return F.linear(h, p["time_embedding.linear_2.weight"], p["time_embedding.linear_2.bias"])


def _sincos_1d(embed_dim: int, pos: np.ndarray) -> np.ndarray:
    # D/models/embeddings.py get_1d_sincos_pos_embed_from_grid (numpy, float64 omega)
    omega = np.arange(embed_dim // 2, dtype=np.float64)
    omega /= embed_dim / 2.0
    omega = 1.0 / 10000**omega
    out = np.einsum("m,d->md", pos.reshape(-1), omega)
    return np.concatenate([np.sin(out), np.cos(out)], axis=1)


def sincos_pos_embed_3d(embed_dim: int, grid_w: int, grid_h: int, frames: int, s_scale: float, t_scale: float) -> np.ndarray:
    """D/models/embeddings.py:81-125 (get_3d_sincos_pos_embed); 2B (non-rotary) models only.
    Returns [frames, grid_h*grid_w, embed_dim]."""
    ds, dt = 3 * embed_dim // 4, embed_dim // 4
    gh = np.arange(grid_h, dtype=np.float32) / s_scale
    gw = np.arange(grid_w, dtype=np.float32) / s_scale
    grid = np.stack(np.meshgrid(gw, gh), axis=0).reshape(2, 1, grid_h, grid_w)
    # get_2d_sincos_pos_embed_from_grid: first half from grid[0] (w), second from grid[1] (h)
    emb_s = np.concatenate([_sincos_1d(ds // 2, grid[0]), _sincos_1d(ds // 2, grid[1])], axis=1)
    gt = np.arange(frames, dtype=np.float32) / t_scale
    emb_t = _sincos_1d(dt, gt)
    emb_s = np.repeat(emb_s[np.newaxis], frames, axis=0)
    emb_t = np.repeat(emb_t[:, np.newaxis], grid_h * grid_w, axis=1)
    return np.concatenate([emb_t, emb_s], axis=-1)


def resize_crop_region_for_grid(src: Tuple[int, int], tgt_width: int, tgt_height: int):
    """D/pipelines/cogvideo/pipeline_cogvideox.py:62-77."""
    tw, th = tgt_width, tgt_height
    h, w = src
    if h / w > th / tw:
        rh, rw = th, int(round(th / h * w))
    else:
        rw, rh = tw, int(round(tw / w * h))
    top, left = int(round((th - rh) / 2.0)), int(round((tw - rw) / 2.0))
    return (top, left), (top + rh, left + rw)


def _rope_1d(dim: int, pos: np.ndarray, theta: float = 10000.0):
    # D/models/embeddings.py:673-736 with use_real=True, repeat_interleave_real=True
    pos_t = torch.from_numpy(pos)
    freqs = 1.0 / (theta ** (torch.arange(0, dim, 2, dtype=torch.float32)[: dim // 2] / dim))
    freqs = torch.outer(pos_t, freqs)
    return freqs.cos().repeat_interleave(2, dim=1).float(), freqs.sin().repeat_interleave(2, dim=1).float()


def rope_3d_tables(head_dim: int, crops, grid_size: Tuple[int, int], temporal_size: int):
    """D/models/embeddings.py:505-570 (get_3d_rotary_pos_embed): cos,sin fp32 [T*h*w, head_dim],
    head_dim split t/h/w = d/4, 3d/8, 3d/8."""
    (s0, s1), (e0, e1) = crops
    gh, gw = grid_size
    grid_h = np.linspace(s0, e0, gh, endpoint=False, dtype=np.float32)
    grid_w = np.linspace(s1, e1, gw, endpoint=False, dtype=np.float32)
    grid_t = np.linspace(0, temporal_size, temporal_size, endpoint=False, dtype=np.float32)
    dt, dh, dw = head_dim // 4, head_dim // 8 * 3, head_dim // 8 * 3
    tc, ts = _rope_1d(dt, grid_t)
    hc, hs = _rope_1d(dh, grid_h)
    wc, ws = _rope_1d(dw, grid_w)

    def combine(t, h, w):
        t = t[:, None, None, :].expand(-1, gh, gw, -1)
        h = h[None, :, None, :].expand(temporal_size, -1, gw, -1)
        w = w[None, None, :, :].expand(temporal_size, gh, -1, -1)
        return torch.cat([t, h, w], dim=-1).reshape(temporal_size * gh * gw, -1)

    return combine(tc, hc, wc), combine(ts, hs, ws)


def pipeline_rope_tables(height: int, width: int, latent_frames: int, head_dim: int = 64, patch: int = 2, vae_sf: int = 8):
    """D/pipelines/cogvideo/pipeline_cogvideox.py:436-460 called with temporal size F+1, then sliced as in
    S/custom_cogvideox_pipe.py:223-235 (reference hard-codes 14 and 1350; generalised per SURVEY §0.7:
    14 -> latent_frames+1, 1350 -> (H/16)(W/16)).  Returns (video_cos, video_sin), (ref_cos, ref_sin)."""
    gh, gw = height // (vae_sf * patch), width // (vae_sf * patch)
    base_w, base_h = 720 // (vae_sf * patch), 480 // (vae_sf * patch)
    crops = resize_crop_region_for_grid((gh, gw), base_w, base_h)
    cos, sin = rope_3d_tables(head_dim, crops, (gh, gw), latent_frames + 1)
    n = gh * gw
    return (cos[n : n * (latent_frames + 1)], sin[n : n * (latent_frames + 1)]), (cos[:n], sin[:n])


def apply_rotary_emb(x: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor) -> torch.Tensor:
    """D/models/embeddings.py:759-778: interleaved-pair rotation in fp32, cast back."""
    xr, xi = x.reshape(*x.shape[:-1], -1, 2).unbind(-1)
    rot = torch.stack([-xi, xr], dim=-1).flatten(3)
    return (x.float() * cos[None, None] + rot.float() * sin[None, None]).to(x.dtype)


# --------------------------------------------------------------------------- block
def layernorm_zero(p: Params, cfg: TransformerConfig, name: str, vid, txt, ref, temb):
    """D/models/normalization.py:467-484 (CogVideoXLayerNormZero.forward).  `enable_lora(..., False)` only sets an
    attribute nobody reads (normalization.py:434-450), so BOTH linear calls run with LoRA applied and
    cond_{shift,scale,gate} == {shift,scale,gate}.  video & ref use chunks 0-2, text 3-5."""
    D = cfg.inner_dim
    mod = lora_linear(p, name + ".linear", F.silu(temb), cfg.lora_scale)
    shift, scale, gate, e_shift, e_scale, e_gate = mod.chunk(6, dim=1)
    w, b = p[name + ".norm.weight"], p[name + ".norm.bias"]

    def ln(x):
        return F.layer_norm(x, (D,), w, b, cfg.norm_eps)

    vid = ln(vid) * (1 + scale)[:, None, :] + shift[:, None, :]
    txt = ln(txt) * (1 + e_scale)[:, None, :] + e_shift[:, None, :]
    ref = ln(ref) * (1 + scale)[:, None, :] + shift[:, None, :]
    return vid, txt, ref, gate[:, None, :], e_gate[:, None, :], gate[:, None, :]


def joint_attention(p: Params, cfg: TransformerConfig, name: str, vid, enc_cat, text_len: int, ref_len: int,
                    rope_video=None, rope_ref=None):
    """D/models/attention_processor.py:2024-2097 (CogVideoXAttnProcessor2_0.__call__): ONE joint self-attention
    over [text | ref | video]; per-head LayerNorm(64, eps 1e-6) on q,k; RoPE on video rows with the video table
    and on ref rows with the ref (frame-0) table; text rows untouched; SDPA scale 1/sqrt(64), no mask."""
    H, d = cfg.num_attention_heads, cfg.attention_head_dim
    x = torch.cat([enc_cat, vid], dim=1)
    B = x.shape[0]
    q = lora_linear(p, name + ".to_q", x, cfg.lora_scale).view(B, -1, H, d).transpose(1, 2)
    k = lora_linear(p, name + ".to_k", x, cfg.lora_scale).view(B, -1, H, d).transpose(1, 2)
    v = lora_linear(p, name + ".to_v", x, cfg.lora_scale).view(B, -1, H, d).transpose(1, 2)
    q = F.layer_norm(q, (d,), p[name + ".norm_q.weight"], p[name + ".norm_q.bias"], 1e-6)
    k = F.layer_norm(k, (d,), p[name + ".norm_k.weight"], p[name + ".norm_k.bias"], 1e-6)
    enc_len = text_len + ref_len
    if rope_video is not None:
        q = q.clone()
        k = k.clone()
        q[:, :, enc_len:] = apply_rotary_emb(q[:, :, enc_len:], *rope_video)
        k[:, :, enc_len:] = apply_rotary_emb(k[:, :, enc_len:], *rope_video)
        # embed_ref_img=True, position_delta=0 (cogvideox_transformer_3d.py:512-515)
        q[:, :, text_len:enc_len] = apply_rotary_emb(q[:, :, text_len:enc_len], *rope_ref) + 0
        k[:, :, text_len:enc_len] = apply_rotary_emb(k[:, :, text_len:enc_len], *rope_ref) + 0
    o = F.scaled_dot_product_attention(q, k, v, dropout_p=0.0, is_causal=False)
    o = o.transpose(1, 2).reshape(B, -1, H * d)
    o = lora_linear(p, name + ".to_out.0", o, cfg.lora_scale)
    return o[:, enc_len:], o[:, :enc_len]


def feed_forward(p: Params, cfg: TransformerConfig, name: str, x):
    """D/models/attention.py:1237-1243 with GELU(tanh) (activations.py:65-90).  `ff.net.0.proj` matches the LoRA
    target suffix "proj", `ff.net.2` is listed explicitly (S/inference.py:222)."""
    h = lora_linear(p, name + ".net.0.proj", x, cfg.lora_scale)
    h = F.gelu(h, approximate="tanh")
    return lora_linear(p, name + ".net.2", h, cfg.lora_scale)


def block_forward(p: Params, cfg: TransformerConfig, prefix: str, vid, txt, temb, ref, rope_video=None, rope_ref=None):
    """D/models/transformers/cogvideox_transformer_3d.py:122-186 (CogVideoXBlock.forward)."""
    text_len, ref_len = txt.shape[1], ref.shape[1]
    n_vid, n_txt, n_ref, g, eg, cg = layernorm_zero(p, cfg, prefix + "norm1", vid, txt, ref, temb)
    a_vid, a_enc = joint_attention(p, cfg, prefix + "attn1", n_vid, torch.cat([n_txt, n_ref], dim=1),
                                   text_len, ref_len, rope_video, rope_ref)
    vid = vid + g * a_vid
    txt = txt + eg * a_enc[:, :text_len]
    ref = ref + cg * a_enc[:, text_len:]
    n_vid, n_txt, n_ref, g, eg, cg = layernorm_zero(p, cfg, prefix + "norm2", vid, txt, ref, temb)
    ff = feed_forward(p, cfg, prefix + "ff", torch.cat([n_txt, n_ref, n_vid], dim=1))
    enc_len = text_len + ref_len
    vid = vid + g * ff[:, enc_len:]
    txt = txt + eg * ff[:, :text_len]
    ref = ref + cg * ff[:, text_len:enc_len]
    return vid, txt, ref


# --------------------------------------------------------------------------- transformer
def patch_embed_image(p: Params, cfg: TransformerConfig, x: torch.Tensor) -> torch.Tensor:
    """Conv2d(k=p, stride=p) patchify of [B,F,C,H,W] -> [B, F*h*w, D] (D/models/embeddings.py:414-419)."""
    B, Fr, C, H, W = x.shape
    y = lora_conv2d(p, "patch_embed.proj", x.reshape(-1, C, H, W), cfg.patch_size, cfg.lora_scale)
    y = y.view(B, Fr, *y.shape[1:]).flatten(3).transpose(2, 3).flatten(1, 2)
    return y


def transformer_forward(p: Params, cfg: TransformerConfig, hidden_states, ref_img_states, encoder_hidden_states,
                        timestep, rope_video=None, rope_ref=None, eval: bool = True):
    """D/models/transformers/cogvideox_transformer_3d.py:450-560.
    hidden_states [B,F,C,H,W]; ref_img_states [Br,1,C,H,W]; encoder_hidden_states [B,L,text_dim]; timestep [B]."""
    dtype = hidden_states.dtype
    B, Fr, C, H, W = hidden_states.shape
    emb = time_embedding(p, cfg, timestep, dtype)
    txt = lora_linear(p, "patch_embed.text_proj", encoder_hidden_states, cfg.lora_scale)        # :494
    ref = patch_embed_image(p, cfg, ref_img_states)                                            # :496-501
    if eval:
        ref = torch.cat([ref, ref], dim=0)                                                     # :503-504
    vid = patch_embed_image(p, cfg, hidden_states)                                             # :506 (text part dropped :516)
    if not cfg.use_rotary_positional_embeddings:
        # 2B: sincos table rebuilt every forward and added to the video tokens only (embeddings.py:433-446;
        # the joint table is zero over the text range).  Ref tokens get NO positional embedding.
        ps = cfg.patch_size
        pe = sincos_pos_embed_3d(cfg.inner_dim, W // ps, H // ps, Fr, cfg.spatial_interpolation_scale,
                                 cfg.temporal_interpolation_scale)
        pe = torch.from_numpy(pe).flatten(0, 1).to(torch.float32)  # copy_ into a float32 zeros table
        vid = vid + pe[None].to(dtype)
    for i in range(cfg.num_layers):
        vid, txt, ref = block_forward(p, cfg, f"transformer_blocks.{i}.", vid, txt, emb, ref, rope_video, rope_ref)
    D = cfg.inner_dim
    vid = F.layer_norm(vid, (D,), p["norm_final.weight"], p["norm_final.bias"], cfg.norm_eps)   # :536-539 (row-wise)
    mod = F.linear(F.silu(emb), p["norm_out.linear.weight"], p["norm_out.linear.bias"])         # normalization.py:70-81
    shift, scale = mod.chunk(2, dim=1)
    vid = F.layer_norm(vid, (D,), p["norm_out.norm.weight"], p["norm_out.norm.bias"], cfg.norm_eps)
    vid = vid * (1 + scale[:, None, :]) + shift[:, None, :]
    out = F.linear(vid, p["proj_out.weight"], p["proj_out.bias"])                               # :543
    ps = cfg.patch_size
    out = out.reshape(B, Fr, H // ps, W // ps, -1, ps, ps)                                      # :549-551
    return out.permute(0, 1, 4, 2, 5, 3, 6).flatten(5, 6).flatten(3, 4)


# --------------------------------------------------------------------------- scheduler (fp64 / integer: bit-exact)
def ddim_alphas_cumprod(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, snr_shift_scale=3.0,
                        rescale_betas_zero_snr=True) -> np.ndarray:
    """D/schedulers/scheduling_ddim_cogvideox.py:203-218 + :95-123, float64.  torch float64 ops are used (not numpy)
    because torch.linspace's two-sided evaluation differs from np.linspace in the last bit and the table must be
    bit-exact."""
    betas = torch.linspace(beta_start**0.5, beta_end**0.5, num_train_timesteps, dtype=torch.float64) ** 2
    ac = torch.cumprod(1.0 - betas, dim=0)
    ac = ac / (snr_shift_scale + (1 - snr_shift_scale) * ac)
    if rescale_betas_zero_snr:
        s = ac.sqrt()
        s0, sT = s[0].clone(), s[-1].clone()
        s = s - sT
        s = s * (s0 / (s0 - sT))
        ac = s**2
    return ac.numpy()


def ddim_timesteps(num_inference_steps: int, num_train_timesteps: int = 1000, spacing: str = "trailing") -> np.ndarray:
    """scheduling_ddim_cogvideox.py:276-301."""
    if spacing == "trailing":
        ratio = num_train_timesteps / num_inference_steps
        return np.round(np.arange(num_train_timesteps, 0, -ratio)).astype(np.int64) - 1
    if spacing == "leading":
        ratio = num_train_timesteps // num_inference_steps
        return (np.arange(0, num_inference_steps) * ratio).round()[::-1].copy().astype(np.int64)
    if spacing == "linspace":
        return np.linspace(0, num_train_timesteps - 1, num_inference_steps).round()[::-1].copy().astype(np.int64)
    raise ValueError(spacing)


def ddim_coefficients(ac: np.ndarray, t: int, num_inference_steps: int, num_train_timesteps: int = 1000):
    """Scalars of scheduling_ddim_cogvideox.py:365-392 (v-prediction), float64:
    returns (sqrt(a_t), sqrt(1-a_t), a_coef, b_coef)."""
    prev = t - num_train_timesteps // num_inference_steps
    a_t = ac[t]
    a_prev = ac[prev] if prev >= 0 else np.float64(1.0)
    beta_t = 1 - a_t
    a_coef = ((1 - a_prev) / (1 - a_t)) ** 0.5
    b_coef = a_prev**0.5 - a_t**0.5 * a_coef
    return a_t**0.5, beta_t**0.5, a_coef, b_coef


def ddim_step(ac: np.ndarray, model_output: torch.Tensor, t: int, sample: torch.Tensor, num_inference_steps: int):
    """scheduling_ddim_cogvideox.py:383-394.  The reference multiplies 0-dim fp64 CPU tensors into the operands, so
    `coef * bf16_tensor` is evaluated with an fp32 scalar and ROUNDED TO bf16 before it meets the fp32 terms."""
    sa, sb, a, b = (torch.tensor(float(c), dtype=torch.float64) for c in ddim_coefficients(ac, t, num_inference_steps))
    x0 = sa * sample - sb * model_output
    return a * sample + b * x0, x0


def dpm_step(ac: np.ndarray, model_output: torch.Tensor, old_x0: Optional[torch.Tensor], t: int, t_back: Optional[int],
             sample: torch.Tensor, num_inference_steps: int, generator=None, num_train_timesteps: int = 1000):
    """CogVideoXDPMScheduler.step (D/schedulers/scheduling_dpm_cogvideox.py:383-439; get_variables :306-317, get_mult :319-328),
    v-prediction, with 0-dim fp64 coefficient tensors like the reference.  Draws noise with the reference's protocol
    (one bf16/fp32 randn per call, a second one on second-order steps).  Returns (prev_sample, pred_original_sample)."""
    act = torch.from_numpy(np.asarray(ac, dtype=np.float64))
    prev_t = t - num_train_timesteps // num_inference_steps
    a_t = act[t]
    a_prev = act[prev_t] if prev_t >= 0 else torch.tensor(1.0)
    a_back = act[t_back] if t_back is not None else None
    x0 = (a_t**0.5) * sample - ((1 - a_t) ** 0.5) * model_output
    lamb = ((a_t / (1 - a_t)) ** 0.5).log()
    lamb_next = ((a_prev / (1 - a_prev)) ** 0.5).log()
    h = lamb_next - lamb
    mult1 = ((1 - a_prev) / (1 - a_t)) ** 0.5 * (-h).exp()
    mult2 = (-2 * h).expm1() * a_prev**0.5
    mult_noise = (1 - a_prev) ** 0.5 * (1 - (-2 * h).exp()) ** 0.5
    noise = torch.randn(sample.shape, generator=generator, dtype=sample.dtype)
    prev = mult1 * sample - mult2 * x0 + mult_noise * noise
    if old_x0 is None or prev_t < 0:
        return prev, x0
    r = (lamb - ((a_back / (1 - a_back)) ** 0.5).log()) / h
    d = (1 + 1 / (2 * r)) * x0 - (1 / (2 * r)) * old_x0
    noise = torch.randn(sample.shape, generator=generator, dtype=sample.dtype)
    return mult1 * sample - mult2 * d + mult_noise * noise, x0


def cfg_combine(noise_pred: torch.Tensor, guidance: float) -> torch.Tensor:
    """S/custom_cogvideox_pipe.py:266,277-279: fp32, uncond first."""
    u, t = noise_pred.float().chunk(2)
    return u + guidance * (t - u)


def dynamic_guidance(guidance_scale: float, i: int, num_inference_steps: int) -> float:
    """S/custom_cogvideox_pipe.py:269-272 — uses the step INDEX i."""
    return 1 + guidance_scale * ((1 - math.cos(math.pi * ((num_inference_steps - i) / num_inference_steps) ** 5.0)) / 2)


def denoise_loop(p: Params, cfg: TransformerConfig, latents, prompt_embeds, ref_img_states, height: int, width: int,
                 num_inference_steps: int = 50, guidance_scale: float = 6.0, use_dynamic_cfg: bool = False,
                 snr_shift_scale: float = 1.0, steps_to_run: Optional[int] = None):
    """S/custom_cogvideox_pipe.py:241-311 with the DDIM scheduler.  `latents` [P,F,16,h,w], `prompt_embeds`
    [2P,L,text_dim] ordered [negative..., positive...] (:196), `ref_img_states` [P,1,16,h,w]."""
    ac = ddim_alphas_cumprod(snr_shift_scale=snr_shift_scale)
    ts = ddim_timesteps(num_inference_steps)
    rope_v = rope_r = None
    if cfg.use_rotary_positional_embeddings:
        rope_v, rope_r = pipeline_rope_tables(height, width, latents.shape[1], cfg.attention_head_dim, cfg.patch_size)
    dtype = prompt_embeds.dtype
    for i, t in enumerate(ts[: steps_to_run or len(ts)]):
        x = torch.cat([latents] * 2)
        tt = torch.full((x.shape[0],), int(t), dtype=torch.int64)
        v = transformer_forward(p, cfg, x, ref_img_states, prompt_embeds.to(x.dtype), tt, rope_v, rope_r, eval=True)
        g = dynamic_guidance(guidance_scale, i, num_inference_steps) if use_dynamic_cfg else guidance_scale
        v = cfg_combine(v, g)
        latents = ddim_step(ac, v, int(t), latents, num_inference_steps)[0].to(dtype)
    return latents


# --------------------------------------------------------------------------- synthetic weights
LORA_LINEAR_SUFFIXES = ("attn1.to_q", "attn1.to_k", "attn1.to_v", "attn1.to_out.0", "norm1.linear", "norm2.linear",
                        "ff.net.0.proj", "ff.net.2")


def param_shapes(cfg: TransformerConfig):
    """Reference state_dict key -> shape, in `state_dict()` order, then the LoRA factors."""
    D, T, ps = cfg.inner_dim, cfg.time_embed_dim, cfg.patch_size
    d = cfg.attention_head_dim
    out = {}
    out["patch_embed.proj.weight"] = (D, cfg.in_channels, ps, ps)
    out["patch_embed.proj.bias"] = (D,)
    out["patch_embed.text_proj.weight"] = (D, cfg.text_embed_dim)
    out["patch_embed.text_proj.bias"] = (D,)
    out["time_embedding.linear_1.weight"] = (T, D)
    out["time_embedding.linear_1.bias"] = (T,)
    out["time_embedding.linear_2.weight"] = (T, T)
    out["time_embedding.linear_2.bias"] = (T,)
    for i in range(cfg.num_layers):
        pre = f"transformer_blocks.{i}."
        for nm in ("norm1", "norm2"):
            out[pre + nm + ".linear.weight"] = (6 * D, T)
            out[pre + nm + ".linear.bias"] = (6 * D,)
            out[pre + nm + ".norm.weight"] = (D,)
            out[pre + nm + ".norm.bias"] = (D,)
        for nm in ("norm_q", "norm_k"):
            out[pre + "attn1." + nm + ".weight"] = (d,)
            out[pre + "attn1." + nm + ".bias"] = (d,)
        for nm in ("to_q", "to_k", "to_v", "to_out.0"):
            out[pre + "attn1." + nm + ".weight"] = (D, D)
            out[pre + "attn1." + nm + ".bias"] = (D,)
        out[pre + "ff.net.0.proj.weight"] = (4 * D, D)
        out[pre + "ff.net.0.proj.bias"] = (4 * D,)
        out[pre + "ff.net.2.weight"] = (D, 4 * D)
        out[pre + "ff.net.2.bias"] = (D,)
    out["norm_final.weight"] = (D,)
    out["norm_final.bias"] = (D,)
    out["norm_out.linear.weight"] = (2 * D, T)
    out["norm_out.linear.bias"] = (2 * D,)
    out["norm_out.norm.weight"] = (D,)
    out["norm_out.norm.bias"] = (D,)
    out["proj_out.weight"] = (ps * ps * cfg.out_channels, D)
    out["proj_out.bias"] = (ps * ps * cfg.out_channels,)
    if cfg.lora_rank:
        r = cfg.lora_rank
        out["patch_embed.proj.lora_A.weight"] = (r, cfg.in_channels, ps, ps)
        out["patch_embed.proj.lora_B.weight"] = (D, r, 1, 1)
        out["patch_embed.text_proj.lora_A.weight"] = (r, cfg.text_embed_dim)
        out["patch_embed.text_proj.lora_B.weight"] = (D, r)
        for i in range(cfg.num_layers):
            pre = f"transformer_blocks.{i}."
            for sfx in LORA_LINEAR_SUFFIXES:
                o, k = out[pre + sfx + ".weight"]
                out[pre + sfx + ".lora_A.weight"] = (r, k)
                out[pre + sfx + ".lora_B.weight"] = (o, r)
    return out


def synth_params(cfg: TransformerConfig, seed: int = 0, dtype=torch.float32, std: float = 0.02) -> Params:
    """Deterministic random weights: Linear/Conv weights & biases N(0, std), LN weight 1+N(0,0.1),
    LN bias N(0,0.1), LoRA A and B N(0, std) (NOT PEFT's zero-init B, or the adapter would be invisible)."""
    g = torch.Generator().manual_seed(seed)
    p = {}
    for k, shp in param_shapes(cfg).items():
        is_ln = (".norm." in k or k.startswith("norm_final") or "norm_q" in k or "norm_k" in k)
        if is_ln and k.endswith("weight"):
            t = 1.0 + 0.1 * torch.randn(shp, generator=g)
        elif is_ln:
            t = 0.1 * torch.randn(shp, generator=g)
        else:
            t = std * torch.randn(shp, generator=g)
        p[k] = t.to(dtype)
    return p
