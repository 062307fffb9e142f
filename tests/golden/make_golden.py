"""Generate the committed golden fixtures by executing the UNMODIFIED reference modules.

Run in the build container only (needs /root/reference):  python tests/golden/make_golden.py
Outputs (tests/golden/*.pt|*.npz) are committed; the GPU box never reads /root/reference.

Every fixture is produced by the reference's own classes:
  CogVideoXDDIMScheduler, get_3d_rotary_pos_embed / CogVideoXPipeline._prepare_rotary_positional_embeddings,
  CogVideoXBlock, CogVideoXTransformer3DModel, CustomCogVideoXPipeline.__call__
with weights from oracle.s2v_oracle.synth_params (seeded) loaded through load_state_dict, and LoRA injected with the
PEFT-layout stand-in of _peft_like.py (peft itself is absent and un-vendored; "parity unpinned" on that arithmetic).
"""
import hashlib
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import _ref_import  # noqa: E402

_ref_import.install()

import _peft_like  # noqa: E402
from oracle import s2v_oracle as O  # noqa: E402

from diffusers.models.embeddings import get_3d_rotary_pos_embed  # noqa: E402
from diffusers.models.transformers.cogvideox_transformer_3d import (  # noqa: E402
    CogVideoXBlock,
    CogVideoXTransformer3DModel,
)
from diffusers.pipelines.cogvideo.pipeline_cogvideox import get_resize_crop_region_for_grid  # noqa: E402
from diffusers.schedulers.scheduling_ddim_cogvideox import CogVideoXDDIMScheduler  # noqa: E402


def sha(t) -> str:
    a = t.detach().contiguous().cpu().numpy() if isinstance(t, torch.Tensor) else np.ascontiguousarray(t)
    return hashlib.sha256(a.tobytes()).hexdigest()


def make_scheduler(snr):
    return CogVideoXDDIMScheduler(
        snr_shift_scale=snr, beta_end=0.012, beta_schedule="scaled_linear", beta_start=0.00085, clip_sample=False,
        num_train_timesteps=1000, prediction_type="v_prediction", rescale_betas_zero_snr=True, set_alpha_to_one=True,
        timestep_spacing="trailing",
    )


# ------------------------------------------------------------------ A. scheduler
def gen_scheduler():
    out = {}
    for tag, snr in (("5b", 1.0), ("2b", 3.0)):
        s = make_scheduler(snr)
        out[f"alphas_cumprod_{tag}"] = s.alphas_cumprod.numpy().astype(np.float64)
        for n in (50, 7, 30):
            s.set_timesteps(n)
            out[f"timesteps_{tag}_{n}"] = s.timesteps.numpy().astype(np.int64)
        # 50-step trace, reference dtype flow: sample bf16, model_output fp32, result fp32 then cast to bf16 by the pipe
        s.set_timesteps(50)
        g = torch.Generator().manual_seed(1234)
        sample = torch.randn(1, 2, 16, 4, 6, generator=g).to(torch.bfloat16)
        out[f"trace_{tag}_sample0"] = sample.float().numpy()
        prevs, x0s, mos = [], [], []
        for t in s.timesteps:
            mo = torch.randn(sample.shape, generator=g, dtype=torch.float32)
            prev, x0 = s.step(mo, t, sample, eta=0.0, return_dict=False)
            assert prev.dtype == torch.float32
            mos.append(mo.numpy()); prevs.append(prev.numpy()); x0s.append(x0.numpy())
            sample = prev.to(torch.bfloat16)
        out[f"trace_{tag}_model_out"] = np.stack(mos)
        out[f"trace_{tag}_prev"] = np.stack(prevs)
        out[f"trace_{tag}_x0"] = np.stack(x0s)
    np.savez_compressed(os.path.join(HERE, "scheduler_ddim.npz"), **out)
    print("scheduler_ddim.npz", {k: v.shape for k, v in out.items() if "trace" not in k})


# ------------------------------------------------------------------ B. RoPE tables
def ref_pipeline_rope(height, width, temporal):
    # exactly CogVideoXPipeline._prepare_rotary_positional_embeddings (pipeline_cogvideox.py:436-460), vae sf 8, patch 2
    gh, gw = height // 16, width // 16
    crops = get_resize_crop_region_for_grid((gh, gw), 720 // 16, 480 // 16)
    return get_3d_rotary_pos_embed(embed_dim=64, crops_coords=crops, grid_size=(gh, gw), temporal_size=temporal)


def gen_rope():
    out = {}
    cos, sin = get_3d_rotary_pos_embed(64, ((0, 0), (8, 8)), (8, 8), 2)
    out["small_cos"], out["small_sin"] = cos.numpy(), sin.numpy()
    for tag, (h, w, T) in {"480x720_T14": (480, 720, 14), "720x1280_T14": (720, 1280, 14), "480x720_T3": (480, 720, 3)}.items():
        cos, sin = ref_pipeline_rope(h, w, T)
        out[f"{tag}_shape"] = np.array(cos.shape)
        out[f"{tag}_cos_sha"] = np.array(sha(cos)); out[f"{tag}_sin_sha"] = np.array(sha(sin))
        rows = np.linspace(0, cos.shape[0] - 1, 97).astype(np.int64)
        out[f"{tag}_rows"] = rows
        out[f"{tag}_cos_rows"] = cos[rows].numpy(); out[f"{tag}_sin_rows"] = sin[rows].numpy()
    for src, tw, th in (((30, 45), 45, 30), ((45, 80), 45, 30), ((60, 60), 45, 30), ((8, 12), 45, 30)):
        out[f"crop_{src[0]}x{src[1]}"] = np.array(get_resize_crop_region_for_grid(src, tw, th)).reshape(-1)
    np.savez_compressed(os.path.join(HERE, "rope.npz"), **out)
    print("rope.npz ok")


# ------------------------------------------------------------------ C/D. block
def block_io(D, H, T, text_len, n_ref, n_vid, seed):
    g = torch.Generator().manual_seed(seed)
    return dict(
        vid=torch.randn(2, n_vid, D, generator=g), txt=torch.randn(2, text_len, D, generator=g),
        ref=torch.randn(2, n_ref, D, generator=g), temb=torch.randn(2, T, generator=g),
    )


def run_ref_block(cfg, params, io, rope, lora):
    blk = CogVideoXBlock(dim=cfg.inner_dim, num_attention_heads=cfg.num_attention_heads,
                         attention_head_dim=cfg.attention_head_dim, time_embed_dim=cfg.time_embed_dim,
                         attention_bias=True, norm_eps=cfg.norm_eps).float().eval()
    if lora:
        _peft_like.inject(blk, cfg.lora_rank, cfg.lora_alpha)
    bp = {k[len("transformer_blocks.0."):]: v for k, v in params.items() if k.startswith("transformer_blocks.0.")}
    _peft_like.load_flat_params(blk, bp)
    rv, rr = rope if rope is not None else (None, None)
    with torch.no_grad():
        return blk(hidden_states=io["vid"], encoder_hidden_states=io["txt"], temb=io["temb"],
                   enc_hidden_states1=io["ref"], image_rotary_emb=rv, embed_ref_img=True,
                   ref_img_seq_start=io["txt"].shape[1], ref_img_seq_end=io["txt"].shape[1] + io["ref"].shape[1],
                   position_delta=0, ref_image_rotary_emb=rr)


def gen_block():
    # C: tiny block, full tensors, LoRA + RoPE / no LoRA no RoPE
    fix = {}
    cos, sin = get_3d_rotary_pos_embed(64, ((0, 0), (8, 8)), (8, 8), 2)
    rope = ((cos[64:], sin[64:]), (cos[:64], sin[:64]))
    for tag, lora, use_rope in (("lora_rope", True, True), ("plain", False, False)):
        cfg = O.TransformerConfig(num_attention_heads=2, num_layers=1, time_embed_dim=64, text_embed_dim=64,
                                  use_rotary_positional_embeddings=use_rope, lora_rank=8 if lora else 0,
                                  lora_alpha=4.0 if lora else 0.0)
        params = O.synth_params(cfg, seed=11)
        io = block_io(128, 2, 64, 226, 64, 64, seed=12)
        vid, txt, ref = run_ref_block(cfg, params, io, rope if use_rope else None, lora)
        fix[tag] = dict(cfg=cfg.__dict__, seed=11, io=io, out=dict(vid=vid, txt=txt, ref=ref),
                        weight_checksum=float(sum(v.double().sum() for v in params.values())))
        if tag == "lora_rope":
            fix[tag]["params"] = params  # RNG canary: committed weights must equal synth_params(cfg, seed) bit for bit
    torch.save(fix, os.path.join(HERE, "block_tiny.pt"))
    # D: cfg-1 shape (BASELINE.json configs[0]): 1 frame 16x16 latent -> 64 video tokens, 64 ref tokens, text 226, B=2,
    # D=1920/H=30 and D=3072/H=48.  Weights come from the seed (too large to commit); outputs are column-subsampled.
    fix = {}
    for tag, H in (("2b", 30), ("5b", 48)):
        for lora, use_rope in ((False, False), (True, True)):
            cfg = O.TransformerConfig(num_attention_heads=H, num_layers=1, use_rotary_positional_embeddings=use_rope,
                                      lora_rank=128 if lora else 0, lora_alpha=64.0 if lora else 0.0)
            params = O.synth_params(cfg, seed=21)
            io = block_io(cfg.inner_dim, H, 512, 226, 64, 64, seed=22)
            vid, txt, ref = run_ref_block(cfg, params, io, rope if use_rope else None, lora)
            wsum = float(sum(v.double().sum() for v in params.values()))
            fix[f"{tag}_{'lora_rope' if lora else 'plain'}"] = dict(
                cfg=cfg.__dict__, weight_checksum=wsum, io_sha={k: sha(v) for k, v in io.items()},
                out=dict(vid=vid[..., ::16].clone(), txt=txt[..., ::16].clone(), ref=ref[..., ::16].clone()),
                out_sum=dict(vid=float(vid.double().sum()), txt=float(txt.double().sum()), ref=float(ref.double().sum())))
    torch.save(fix, os.path.join(HERE, "block_cfg1.pt"))
    print("block fixtures ok")


# ------------------------------------------------------------------ E. tiny transformer
def build_ref_transformer(cfg, params, lora, dtype=torch.float32):
    m = CogVideoXTransformer3DModel(
        num_attention_heads=cfg.num_attention_heads, attention_head_dim=cfg.attention_head_dim,
        in_channels=cfg.in_channels, out_channels=cfg.out_channels, time_embed_dim=cfg.time_embed_dim,
        text_embed_dim=cfg.text_embed_dim, num_layers=cfg.num_layers, patch_size=cfg.patch_size,
        use_rotary_positional_embeddings=cfg.use_rotary_positional_embeddings).float().eval()
    if lora:
        _peft_like.inject(m, cfg.lora_rank, cfg.lora_alpha)
    _peft_like.load_flat_params(m, params)
    return m.to(dtype)


def gen_transformer():
    fix = {}
    for tag, lora, use_rope in (("lora_rope", True, True), ("plain_sincos", False, False)):
        cfg = O.TransformerConfig(num_attention_heads=2, num_layers=2, time_embed_dim=64, text_embed_dim=64,
                                  use_rotary_positional_embeddings=use_rope, lora_rank=8 if lora else 0,
                                  lora_alpha=4.0 if lora else 0.0)
        params = O.synth_params(cfg, seed=31)
        g = torch.Generator().manual_seed(32)
        Fr, h, w = 3, 8, 12
        io = dict(hidden=torch.randn(2, Fr, 16, h, w, generator=g), ref=0.7 * torch.randn(1, 1, 16, h, w, generator=g),
                  text=0.2 * torch.randn(2, 226, 64, generator=g), timestep=torch.tensor([979, 979]))
        rope = None
        if use_rope:
            cos, sin = ref_pipeline_rope(h * 8, w * 8, Fr + 1)
            n = (h // 2) * (w // 2)
            rope = ((cos[n:], sin[n:]), (cos[:n], sin[:n]))
        m = build_ref_transformer(cfg, params, lora)
        with torch.no_grad():
            out = m(hidden_states=io["hidden"], ref_img_states=io["ref"], encoder_hidden_states=io["text"],
                    timestep=io["timestep"], image_rotary_emb=rope[0] if rope else None,
                    ref_image_rotary_emb=rope[1] if rope else None, return_dict=False, eval=True)[0]
        fix[tag] = dict(cfg=cfg.__dict__, seed=31, io=io, out=out,
                        weight_checksum=float(sum(v.double().sum() for v in params.values())))
    torch.save(fix, os.path.join(HERE, "transformer_tiny.pt"))
    print("transformer_tiny.pt ok")


# ------------------------------------------------------------------ F. the custom pipeline loop
def gen_pipe_loop():
    import transformers.utils  # noqa: F401  (FLAX_WEIGHTS_NAME shim already applied)
    from custom_cogvideox_pipe import CustomCogVideoXPipeline

    cfg = O.TransformerConfig(num_attention_heads=2, num_layers=2, time_embed_dim=64, text_embed_dim=64,
                              use_rotary_positional_embeddings=True, lora_rank=8, lora_alpha=4.0)
    params = O.synth_params(cfg, seed=41)
    fix = dict(cfg=cfg.__dict__, seed=41, runs={},
               weight_checksum=float(sum(v.double().sum() for v in params.values())))
    g = torch.Generator().manual_seed(42)
    P, Fr, h, w = 1, 2, 60, 90  # the reference hard-codes 1350 tokens/frame => 480x720 only
    latents = torch.randn(P, Fr, 16, h, w, generator=g)
    pos = 0.2 * torch.randn(P, 226, 64, generator=g)
    neg = 0.2 * torch.randn(P, 226, 64, generator=g)
    ref = 0.7 * torch.randn(P, 1, 16, h, w, generator=g)
    fix["io"] = dict(latents=latents, prompt_embeds=pos, negative_prompt_embeds=neg, ref_img_states=ref)
    for tag, dtype, dyn in (("fp32", torch.float32, False), ("fp32_dyncfg", torch.float32, True),
                            ("bf16", torch.bfloat16, False)):
        m = build_ref_transformer(cfg, params, True, dtype)
        # The real constructor; T5 / tokenizer / VAE are not on the loop when prompt_embeds and latents are given
        # and output_type="latent" (vae None => scale factors fall back to 8 / 4 / 0.7, pipeline_cogvideox.py:184-192).
        pipe = CustomCogVideoXPipeline(tokenizer=None, text_encoder=None, transformer=m, vae=None,
                                       scheduler=make_scheduler(1.0))
        pipe.set_progress_bar_config(disable=True)
        with torch.no_grad():
            out = pipe(prompt=None, ref_img_states=ref.to(dtype), height=480, width=720, num_frames=(Fr - 1) * 4 + 1,
                       num_inference_steps=3, guidance_scale=6.0, use_dynamic_cfg=dyn, latents=latents.to(dtype),
                       prompt_embeds=pos.to(dtype), negative_prompt_embeds=neg.to(dtype), output_type="latent",
                       return_dict=False, eval=True)[0]
        fix["runs"][tag] = out.float()
        print("pipe loop", tag, out.dtype, float(out.float().abs().mean()))
    torch.save(fix, os.path.join(HERE, "pipe_loop_tiny.pt"))


# ------------------------------------------------------------------ H. DPM scheduler trace (SURVEY §8f "next")
def gen_dpm():
    """50-step trace of the reference CogVideoXDPMScheduler on the CPU (bf16 sample, fp32 model output, seeded CPU generator),
    with the loop protocol of S/custom_cogvideox_pipe.py:288-296 (old_pred_original_sample, timesteps[i-1], .to(bf16))."""
    from diffusers.schedulers.scheduling_dpm_cogvideox import CogVideoXDPMScheduler

    out = {}
    for tag, snr in (("5b", 1.0), ("2b", 3.0)):
        s = CogVideoXDPMScheduler(snr_shift_scale=snr, beta_end=0.012, beta_schedule="scaled_linear", beta_start=0.00085,
                                  clip_sample=False, num_train_timesteps=1000, prediction_type="v_prediction",
                                  rescale_betas_zero_snr=True, set_alpha_to_one=True, timestep_spacing="trailing")
        s.set_timesteps(50)
        g = torch.Generator().manual_seed(4321)
        gn = torch.Generator().manual_seed(99)
        sample = torch.randn(1, 2, 16, 4, 6, generator=g).to(torch.bfloat16)
        out[f"dpm_{tag}_sample0"] = sample.float().numpy()
        old = None
        prevs, x0s, mos = [], [], []
        ts = s.timesteps
        for i, t in enumerate(ts):
            mo = torch.randn(sample.shape, generator=g, dtype=torch.float32)
            prev, old = s.step(mo, old, t, ts[i - 1] if i > 0 else None, sample, generator=gn, return_dict=False)
            assert prev.dtype == torch.float32 and old.dtype == torch.float32
            mos.append(mo.numpy()); prevs.append(prev.numpy()); x0s.append(old.numpy())
            sample = prev.to(torch.bfloat16)
        out[f"dpm_{tag}_model_out"] = np.stack(mos)
        out[f"dpm_{tag}_prev"] = np.stack(prevs)
        out[f"dpm_{tag}_x0"] = np.stack(x0s)
    np.savez_compressed(os.path.join(HERE, "scheduler_dpm.npz"), **out)
    print("scheduler_dpm.npz", float(np.abs(out["dpm_5b_prev"][-1]).mean()))


# ------------------------------------------------------------------ G. VAE decoder (row V)
def gen_vae():
    """The reference AutoencoderKLCogVideoX (decoder half) with seeded weights: one decoder call with a conv-cache chain,
    the untiled `_decode`, and the tiled + blended `decode` (tiling and slicing enabled as in S/inference.py:206-207)."""
    from diffusers.models.autoencoders.autoencoder_kl_cogvideox import AutoencoderKLCogVideoX

    from oracle import vae_oracle as V

    cfg = V.VaeConfig(block_out_channels=(64, 128, 128, 128), layers_per_block=1, sample_height=64, sample_width=96,
                      scaling_factor=0.7)
    params = V.synth_decoder_params(cfg, seed=51)
    vae = AutoencoderKLCogVideoX(block_out_channels=cfg.block_out_channels, layers_per_block=cfg.layers_per_block,
                                 latent_channels=16, sample_height=cfg.sample_height, sample_width=cfg.sample_width,
                                 scaling_factor=cfg.scaling_factor, temporal_compression_ratio=4).float().eval()
    missing, unexpected = vae.load_state_dict(params, strict=False)
    assert not unexpected and all(k.startswith("encoder.") for k in missing), (unexpected, [k for k in missing if not k.startswith("encoder.")][:5])
    assert set(params) == {k for k in vae.state_dict() if k.startswith("decoder.")}
    g = torch.Generator().manual_seed(52)
    z = torch.randn(1, 16, 5, 8, 12, generator=g)            # 5 latent frames (batches [0:3] [3:5]), 8x12 latent = 64x96 px
    fix = dict(cfg=cfg.__dict__, seed=51, z=z, weight_checksum=float(sum(v.double().sum() for v in params.values())))
    with torch.no_grad():
        # one tile-sized decoder call chain with caches (4x6 latent tile)
        zt = z[:, :, :, :4, :6]
        y0, cache = vae.decoder(zt[:, :, :3], conv_cache=None)
        y1, _ = vae.decoder(zt[:, :, 3:5], conv_cache=cache)
        fix["decoder_chain"] = dict(y0=y0, y1=y1)
        fix["decode_untiled"] = vae.decode(z).sample                     # tiling off
        vae.enable_slicing()
        vae.enable_tiling()
        fix["decode_tiled"] = vae.decode(z).sample                       # 3x3 tiles, blended
        z2 = torch.cat([z, 0.5 * z.flip(3)], dim=0)
        fix["decode_tiled_b2_sum"] = float(vae.decode(z2).sample.double().sum())
    print("vae", {k: tuple(v.shape) for k, v in fix.items() if isinstance(v, torch.Tensor)}, fix["decode_tiled"].abs().mean())
    torch.save(fix, os.path.join(HERE, "vae_tiny.pt"))


# ------------------------------------------------------------------ I. VAE encoder, reference-image path
def gen_vae_enc():
    """The reference AutoencoderKLCogVideoX.encode on ONE frame (S/video_generate.py:26-38) with seeded encoder weights: untiled,
    and tiled (3x3 overlapping sample tiles, blended in latent space) as S/inference.py:206-207 enables it; plus the
    reference-image preparation chain end to end with a fixed noise draw."""
    from diffusers.models.autoencoders.autoencoder_kl_cogvideox import AutoencoderKLCogVideoX

    from oracle import vae_oracle as V

    cfg = V.VaeConfig(block_out_channels=(64, 128, 128, 128), layers_per_block=1, sample_height=64, sample_width=96,
                      scaling_factor=0.7)
    params = V.synth_encoder_params(cfg, seed=61)
    vae = AutoencoderKLCogVideoX(block_out_channels=cfg.block_out_channels, layers_per_block=cfg.layers_per_block,
                                 latent_channels=16, sample_height=cfg.sample_height, sample_width=cfg.sample_width,
                                 scaling_factor=cfg.scaling_factor, temporal_compression_ratio=4).float().eval()
    missing, unexpected = vae.load_state_dict(params, strict=False)
    assert not unexpected and all(k.startswith("decoder.") for k in missing), (unexpected, [k for k in missing if not k.startswith("decoder.")][:5])
    assert set(params) == {k for k in vae.state_dict() if k.startswith("encoder.")}
    g = torch.Generator().manual_seed(62)
    img = torch.randint(0, 256, (64, 96, 3), generator=g, dtype=torch.uint8).numpy()
    x = torch.from_numpy(np.expand_dims(img, 0)).float() / 255.0 * 2.0 - 1.0                   # S/video_generate.py:29-33
    x = x.permute(0, 3, 1, 2).unsqueeze(0).permute(0, 2, 1, 3, 4)
    fix = dict(cfg=cfg.__dict__, seed=61, image=img, weight_checksum=float(sum(v.double().sum() for v in params.values())))
    with torch.no_grad():
        fix["moments_untiled"] = vae.encode(x).latent_dist.parameters
        fix["moments_tile"] = vae.encode(x[:, :, :, :32, :48]).latent_dist.parameters        # one tile-sized call
        vae.enable_slicing()
        vae.enable_tiling()
        dist = vae.encode(x).latent_dist
        fix["moments_tiled"] = dist.parameters
        gen = torch.Generator().manual_seed(63)
        fix["ref_img_states"] = (dist.sample(gen) * vae.config.scaling_factor).permute(0, 2, 1, 3, 4)   # S/video_generate.py:36-38
        fix["noise"] = torch.randn(dist.mean.shape, generator=torch.Generator().manual_seed(63))
    print("vae_enc", {k: tuple(v.shape) for k, v in fix.items() if isinstance(v, torch.Tensor)}, float(fix["moments_tiled"].abs().mean()))
    torch.save(fix, os.path.join(HERE, "vae_enc_tiny.pt"))


# ------------------------------------------------------------------ H. host glue after the decoder
def gen_post():
    """The reference's own VideoProcessor.postprocess_video (np and pil outputs) and export_to_video's uint8 conversion on a
    seeded bf16 video that includes out-of-range values and exact rounding ties."""
    import numpy as np
    from diffusers.video_processor import VideoProcessor

    g = torch.Generator().manual_seed(71)
    v = (torch.randn(2, 3, 3, 8, 12, generator=g) * 0.8)
    ties = torch.tensor([-1.0, 1.0, 0.0, -0.5, 0.5, 1.5, -1.5, 1 / 255.0 - 1.0, 0.00390625, 0.99609375, -0.99609375, 0.5 / 255.0])
    v.view(-1)[: ties.numel()] = ties
    v = v.to(torch.bfloat16)
    vp = VideoProcessor(vae_scale_factor=8)
    as_np = vp.postprocess_video(video=v, output_type="np")                     # [B,F,H,W,C] fp32 in [0,1]
    trunc = np.stack([np.stack([(frame * 255).astype(np.uint8) for frame in vid]) for vid in as_np])   # export_utils.py:177-178
    pil = vp.postprocess_video(video=v, output_type="pil")
    rounded = np.stack([np.stack([np.array(im) for im in vid]) for vid in pil])
    np.savez_compressed(os.path.join(HERE, "postprocess.npz"), video_bf16_bits=v.view(torch.int16).numpy(), as_np=as_np, trunc=trunc,
                        rounded=rounded)
    print("postprocess.npz", trunc.shape, int((trunc != rounded).sum()), "pixels differ between the two roundings")


if __name__ == "__main__":
    torch.manual_seed(0)
    which = sys.argv[1:] or ["sched", "rope", "block", "transformer", "pipe", "vae", "dpm"]
    if "sched" in which: gen_scheduler()
    if "rope" in which: gen_rope()
    if "block" in which: gen_block()
    if "transformer" in which: gen_transformer()
    if "pipe" in which: gen_pipe_loop()
    if "vae" in which: gen_vae()
    if "dpm" in which: gen_dpm()
    if "post" in which: gen_post()
    if "vae_enc" in which: gen_vae_enc()
