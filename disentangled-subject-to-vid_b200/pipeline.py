"""CustomCogVideoXPipeline — the reference's pipeline surface (S/custom_cogvideox_pipe.py:21-326) on the B200 engine.

`__call__` keeps the reference's argument list, defaults, error behaviour and output type, so
S/video_generate.py:47-61 (`pipe(prompt=..., ref_img_states=..., guidance_scale=..., use_dynamic_cfg=..., height=...,
width=..., num_frames=..., num_inference_steps=50, output_type="np", eval=True)`) drops in unchanged.

What differs underneath (same results, different schedule):
  * one fused token buffer and ~12 kernels per block instead of ~60 library calls;
  * CFG combine + DDIM update are ONE bit-exact kernel with host-computed fp64 coefficients — no per-step host sync
    (the reference indexes a CPU table with a CUDA timestep twice per step);
  * RoPE tables are generalised from the hard-coded 14 x 1350 to (latent_frames+1) x (H/16)(W/16);
  * prompts x CFG halves can be sharded over the GPUs of one node (`parallel.py`).
The T5 text encoder and the 3-D VAE are pluggable collaborators (any object with the HF / diffusers call surface);
they are outside the timed denoising loop.
"""
from __future__ import annotations

import math
from typing import Any, Callable, Dict, List, Optional, Tuple, Union

import torch

from . import ops, tables
from .scheduler import CogVideoXDPMScheduler  # noqa: F401
from .scheduler import CogVideoXDDIMScheduler

BF16 = torch.bfloat16


class CogVideoXPipelineOutput:
    def __init__(self, frames):
        self.frames = frames


class CustomCogVideoXPipeline:
    _callback_tensor_inputs = ["latents", "prompt_embeds", "negative_prompt_embeds"]

    def __init__(self, tokenizer, text_encoder, transformer, vae, scheduler, clip_tokenizer=None, clip_text_encoder=None,
                 customization=False, vae_add=False, cross_attend=False, cross_attend_text=False, pos_embed=False,
                 qk_replace=False):
        self.tokenizer, self.text_encoder, self.transformer, self.vae, self.scheduler = tokenizer, text_encoder, transformer, vae, scheduler
        self.customization = customization
        vcfg = getattr(vae, "config", None)
        g = lambda k, d: (vcfg[k] if isinstance(vcfg, dict) else getattr(vcfg, k, d)) if vcfg is not None else d  # noqa: E731
        self.vae_scale_factor_spatial = 2 ** (len(g("block_out_channels", [0] * 4)) - 1)
        self.vae_scale_factor_temporal = g("temporal_compression_ratio", 4)
        self.vae_scaling_factor_image = g("scaling_factor", 0.7)
        self._guidance_scale = 1.0
        self._num_timesteps = 0
        self.interrupt = False
        self._rope_cache: Dict[Tuple, Tuple[torch.Tensor, torch.Tensor]] = {}
        self._cfg_plan = None      # parallel.ShardPlan(mode="cfg") when this rank runs one CFG half (enable_cfg_parallel)
        self._cfg_xchg = None

    # ------------------------------------------------------------------ helpers (base pipeline surface)
    @property
    def _execution_device(self) -> torch.device:
        return next(self.transformer.parameters()).device

    @property
    def guidance_scale(self):
        return self._guidance_scale

    @property
    def num_timesteps(self):
        return self._num_timesteps

    def to(self, device):
        self.transformer.to(device)
        if hasattr(self.transformer, "invalidate_engine"):
            self.transformer.invalidate_engine()
        for m in (self.text_encoder, self.vae):
            if m is not None and hasattr(m, "to"):
                m.to(device)
        return self

    def check_inputs(self, prompt, height, width, negative_prompt, callback_on_step_end_tensor_inputs, prompt_embeds=None,
                     negative_prompt_embeds=None):
        # D/pipelines/cogvideo/pipeline_cogvideox.py:372-421 (same conditions, same exception types)
        if height % 8 != 0 or width % 8 != 0:
            raise ValueError(f"`height` and `width` have to be divisible by 8 but are {height} and {width}.")
        if callback_on_step_end_tensor_inputs is not None and not all(
                k in self._callback_tensor_inputs for k in callback_on_step_end_tensor_inputs):
            bad = [k for k in callback_on_step_end_tensor_inputs if k not in self._callback_tensor_inputs]
            raise ValueError(f"`callback_on_step_end_tensor_inputs` has to be in {self._callback_tensor_inputs}, but found {bad}")
        if prompt is not None and prompt_embeds is not None:
            raise ValueError(f"Cannot forward both `prompt`: {prompt} and `prompt_embeds`: {prompt_embeds}. Please make sure to"
                             " only forward one of the two.")
        elif prompt is None and prompt_embeds is None:
            raise ValueError("Provide either `prompt` or `prompt_embeds`. Cannot leave both `prompt` and `prompt_embeds` undefined.")
        elif prompt is not None and (not isinstance(prompt, str) and not isinstance(prompt, list)):
            raise ValueError(f"`prompt` has to be of type `str` or `list` but is {type(prompt)}")
        if prompt is not None and negative_prompt_embeds is not None:
            raise ValueError(f"Cannot forward both `prompt`: {prompt} and `negative_prompt_embeds`: {negative_prompt_embeds}."
                             " Please make sure to only forward one of the two.")
        if negative_prompt is not None and negative_prompt_embeds is not None:
            raise ValueError(f"Cannot forward both `negative_prompt`: {negative_prompt} and `negative_prompt_embeds`:"
                             f" {negative_prompt_embeds}. Please make sure to only forward one of the two.")
        if prompt_embeds is not None and negative_prompt_embeds is not None and prompt_embeds.shape != negative_prompt_embeds.shape:
            raise ValueError("`prompt_embeds` and `negative_prompt_embeds` must have the same shape when passed directly, but"
                             f" got: `prompt_embeds` {prompt_embeds.shape} != `negative_prompt_embeds` {negative_prompt_embeds.shape}.")

    def _get_t5_prompt_embeds(self, prompt, num_videos_per_prompt=1, max_sequence_length=226, device=None, dtype=None):
        # D/pipelines/cogvideo/pipeline_cogvideox.py:197-237 — needs the pluggable tokenizer / T5 encoder
        if self.tokenizer is None or self.text_encoder is None:
            raise RuntimeError("no tokenizer / text_encoder attached: pass prompt_embeds and negative_prompt_embeds")
        device = device or self._execution_device
        dtype = dtype or next(self.text_encoder.parameters()).dtype
        prompt = [prompt] if isinstance(prompt, str) else prompt
        ids = self.tokenizer(prompt, padding="max_length", max_length=max_sequence_length, truncation=True,
                             add_special_tokens=True, return_tensors="pt").input_ids
        emb = self.text_encoder(ids.to(device))[0].to(dtype=dtype, device=device)
        _, seq, _ = emb.shape
        emb = emb.repeat(1, num_videos_per_prompt, 1)
        return emb.view(len(prompt) * num_videos_per_prompt, seq, -1)

    def encode_prompt(self, prompt, negative_prompt=None, do_classifier_free_guidance=True, num_videos_per_prompt=1,
                      prompt_embeds=None, negative_prompt_embeds=None, negative_clip_prompt_embeds=None,
                      max_sequence_length=226, device=None, dtype=None):
        # S/custom_cogvideox_pipe.py:43-123
        prompt = [prompt] if isinstance(prompt, str) else prompt
        batch_size = len(prompt) if prompt is not None else prompt_embeds.shape[0]
        if prompt_embeds is None:
            prompt_embeds = self._get_t5_prompt_embeds(prompt, num_videos_per_prompt, max_sequence_length, device, dtype)
        if do_classifier_free_guidance and negative_prompt_embeds is None:
            negative_prompt = negative_prompt or ""
            negative_prompt = batch_size * [negative_prompt] if isinstance(negative_prompt, str) else negative_prompt
            if prompt is not None and type(prompt) is not type(negative_prompt):
                raise TypeError(f"`negative_prompt` should be the same type to `prompt`, but got {type(negative_prompt)} !="
                                f" {type(prompt)}.")
            elif batch_size != len(negative_prompt):
                raise ValueError(f"`negative_prompt`: {negative_prompt} has batch size {len(negative_prompt)}, but `prompt`:"
                                 f" {prompt} has batch size {batch_size}. Please make sure that passed `negative_prompt` matches"
                                 " the batch size of `prompt`.")
            negative_prompt_embeds = self._get_t5_prompt_embeds(negative_prompt, num_videos_per_prompt, max_sequence_length, device, dtype)
        return prompt_embeds, negative_prompt_embeds, None, None

    def prepare_latents(self, batch_size, num_channels_latents, num_frames, height, width, dtype, device, generator, latents=None):
        # D/pipelines/cogvideo/pipeline_cogvideox.py:320-344 (randn_tensor: D/utils/torch_utils.py:38-83)
        if isinstance(generator, list) and len(generator) != batch_size:
            raise ValueError(f"You have passed a list of generators of length {len(generator)}, but requested an effective batch"
                             f" size of {batch_size}. Make sure the batch size matches the length of the generators.")
        shape = (batch_size, (num_frames - 1) // self.vae_scale_factor_temporal + 1, num_channels_latents,
                 height // self.vae_scale_factor_spatial, width // self.vae_scale_factor_spatial)
        if latents is None:
            if isinstance(generator, list):
                latents = torch.cat([torch.randn((1,) + shape[1:], generator=g, device=g.device, dtype=dtype) for g in generator]).to(device)
            else:
                rdev = generator.device if generator is not None else device
                latents = torch.randn(shape, generator=generator, device=rdev, dtype=dtype).to(device)
        else:
            latents = latents.to(device)
        return latents * self.scheduler.init_noise_sigma

    def rotary_tables(self, height: int, width: int, latent_frames: int, device) -> Tuple[torch.Tensor, torch.Tensor]:
        """Joint [ref | video] cos/sin table (pipeline_cogvideox.py:436-460 + custom_cogvideox_pipe.py:223-235)."""
        cfg = self.transformer.config
        g = lambda k, d: cfg[k] if isinstance(cfg, dict) else getattr(cfg, k, d)  # noqa: E731
        key = (height, width, latent_frames, g("attention_head_dim", 64), g("patch_size", 2), self.vae_scale_factor_spatial, str(device))
        if key not in self._rope_cache:   # the reference rebuilds the table on the host every call; it only depends on the geometry
            cos, sin = tables.joint_rope_table(*key[:6])
            if len(self._rope_cache) >= 4:
                self._rope_cache.clear()
            self._rope_cache[key] = (cos.to(device), sin.to(device))
        return self._rope_cache[key]

    # ------------------------------------------------------------------ CFG halves on two GPUs (parallel.py)
    def enable_cfg_parallel(self, shard_plan, record_times: bool = False):
        """This rank runs ONE classifier-free-guidance half (shard_plan.cfg_half: 0 = negative prompt, 1 = prompt) of the prompts
        it is called with; its pair rank runs the other half.  Per step the pair exchanges the model output (one NCCL
        all_gather_into_tensor on a side stream, parallel.PairExchange) and both apply the identical bit-exact CFG + scheduler
        kernel, so `__call__` returns the same latents on both ranks — bit-identical to the single-GPU call of the same build.
        Collective: every rank of the job must call this (it creates the pair process groups)."""
        from .parallel import PairExchange
        if shard_plan.mode != "cfg":
            raise ValueError("enable_cfg_parallel needs a ShardPlan in 'cfg' mode (parallel.plan(..., mode='cfg'))")
        self._cfg_plan = shard_plan
        self._cfg_xchg = PairExchange(shard_plan, record_times=record_times)
        return self

    def disable_cfg_parallel(self):
        self._cfg_plan = self._cfg_xchg = None

    def decode_latents(self, latents: torch.Tensor) -> torch.Tensor:
        if self.vae is None:
            raise RuntimeError("no VAE attached: call with output_type='latent'")
        latents = latents.permute(0, 2, 1, 3, 4)
        latents = 1 / self.vae_scaling_factor_image * latents
        return self.vae.decode(latents).sample

    def _guidance_for_step(self, guidance_scale: float, use_dynamic_cfg: bool, i: int, num_inference_steps: int) -> float:
        if use_dynamic_cfg:  # S/custom_cogvideox_pipe.py:269-272 — step INDEX, not timestep
            return 1 + guidance_scale * ((1 - math.cos(math.pi * ((num_inference_steps - i) / num_inference_steps) ** 5.0)) / 2)
        return guidance_scale

    # ------------------------------------------------------------------ the call
    @torch.no_grad()
    def __call__(self, prompt: Optional[Union[str, List[str]]] = None, ref_img_states: Optional[torch.Tensor] = None,
                 negative_prompt: Optional[Union[str, List[str]]] = None, height: int = 480, width: int = 720,
                 num_frames: int = 49, num_inference_steps: int = 50, timesteps: Optional[List[int]] = None,
                 guidance_scale: float = 6, use_dynamic_cfg: bool = False, num_videos_per_prompt: int = 1, eta: float = 0.0,
                 generator=None, latents: Optional[torch.Tensor] = None, prompt_embeds: Optional[torch.Tensor] = None,
                 negative_prompt_embeds: Optional[torch.Tensor] = None, output_type: str = "pil", return_dict: bool = True,
                 attention_kwargs: Optional[Dict[str, Any]] = None,
                 callback_on_step_end: Optional[Callable[[int, int, Dict], None]] = None,
                 callback_on_step_end_tensor_inputs: List[str] = ["latents"], max_sequence_length: int = 226, eval: bool = True):
        if num_frames > 49:
            raise ValueError("The number of frames must be less than or equal to 49 due to static positional embeddings.")
        self.check_inputs(prompt, height, width, negative_prompt, callback_on_step_end_tensor_inputs, prompt_embeds,
                          negative_prompt_embeds)
        if eta != 0.0:
            raise NotImplementedError("the CogVideoX DDIM step ignores eta (scheduling_ddim_cogvideox.py:305-402); eta must be 0")
        device = self._execution_device
        if prompt is not None and isinstance(prompt, str):
            batch_size = 1
        elif prompt is not None and isinstance(prompt, list):
            batch_size = len(prompt)
        else:
            batch_size = prompt_embeds.shape[0]
        do_cfg = guidance_scale > 1.0
        if not do_cfg:
            # eval=True doubles the reference tokens along batch (cogvideox_transformer_3d.py:503-504), so the reference
            # itself only runs with CFG on
            raise ValueError("the subject-to-video transformer requires classifier-free guidance (guidance_scale > 1)")
        prompt_embeds, negative_prompt_embeds, _, _ = self.encode_prompt(
            prompt, negative_prompt=negative_prompt, do_classifier_free_guidance=do_cfg, num_videos_per_prompt=num_videos_per_prompt,
            prompt_embeds=prompt_embeds, negative_prompt_embeds=negative_prompt_embeds, max_sequence_length=max_sequence_length,
            device=device)
        if self._cfg_plan is None:
            prompt_embeds = torch.cat([negative_prompt_embeds, prompt_embeds], dim=0).to(device=device, dtype=BF16)
        else:   # CFG halves on two GPUs: this rank's half of the [negative | positive] batch of S/custom_cogvideox_pipe.py:196
            prompt_embeds = (negative_prompt_embeds if self._cfg_plan.cfg_half == 0 else prompt_embeds).to(device=device, dtype=BF16)

        if timesteps is None:
            self.scheduler.set_timesteps(num_inference_steps, device=device)
            steps_host = list(self.scheduler._timesteps_host)
        else:
            steps_host = [int(t) for t in timesteps]
            self.scheduler.timesteps = torch.tensor(steps_host, device=device)
            self.scheduler._timesteps_host = steps_host
            if self.scheduler.num_inference_steps is None:
                self.scheduler.num_inference_steps = num_inference_steps
        self._num_timesteps = len(steps_host)

        cfg = self.transformer.config
        in_ch = cfg["in_channels"] if isinstance(cfg, dict) else cfg.in_channels
        latents = self.prepare_latents(batch_size * num_videos_per_prompt, in_ch, num_frames, height, width, BF16, device,
                                       generator, latents).to(BF16).contiguous()
        P, Fr = latents.shape[0], latents.shape[1]
        if ref_img_states is None:
            raise ValueError("ref_img_states must be provided: the transformer patch-embeds the reference image every step")
        ref = ref_img_states.to(device=device, dtype=BF16)
        if ref.shape[0] == 1 and P > 1:
            ref = ref.expand(P, *ref.shape[1:])
        ref = ref.contiguous()
        rotary = bool(cfg["use_rotary_positional_embeddings"] if isinstance(cfg, dict) else cfg.use_rotary_positional_embeddings)
        rope = self.rotary_tables(height, width, Fr, device) if rotary else None
        n = rope[0].shape[0] // (Fr + 1) if rotary else 0
        image_rotary_emb = (rope[0][n:], rope[1][n:]) if rotary else None
        ref_image_rotary_emb = (rope[0][:n], rope[1][:n]) if rotary else None

        # timestep tensors for every step, built once (no per-step host->device traffic)
        t_dev = torch.tensor(steps_host, device=device, dtype=torch.float32)
        model_in = torch.empty((2 * P,) + tuple(latents.shape[1:]), device=device, dtype=BF16)
        nxt = torch.empty_like(latents)
        old_pred_original_sample = None
        for i, t in enumerate(steps_host):
            if self.interrupt:
                break
            if self._cfg_plan is None:
                model_in[:P].copy_(latents)     # torch.cat([latents] * 2); scale_model_input is the identity
                model_in[P:].copy_(latents)
                noise_pred = self.transformer(hidden_states=model_in, encoder_hidden_states=prompt_embeds, ref_img_states=ref,
                                              timestep=t_dev[i:i + 1].expand(2 * P), image_rotary_emb=image_rotary_emb,
                                              ref_image_rotary_emb=ref_image_rotary_emb, attention_kwargs=attention_kwargs,
                                              return_dict=False, eval=True)[0]
            else:
                # one CFG half here, the other on the pair rank (eval=False: the reference tokens are not doubled along batch
                # because the batch is not); then (uncond, cond) are exchanged inside the pair
                half = self.transformer(hidden_states=latents, encoder_hidden_states=prompt_embeds, ref_img_states=ref,
                                        timestep=t_dev[i:i + 1].expand(P), image_rotary_emb=image_rotary_emb,
                                        ref_image_rotary_emb=ref_image_rotary_emb, attention_kwargs=attention_kwargs,
                                        return_dict=False, eval=False)[0]
                noise_pred = self._cfg_xchg.both_halves(half)
            self._guidance_scale = self._guidance_for_step(guidance_scale, use_dynamic_cfg, i, num_inference_steps)
            # .float() -> u + g (t - u) -> scheduler step -> .to(bf16), fused and bit-exact
            if isinstance(self.scheduler, CogVideoXDPMScheduler):   # S/custom_cogvideox_pipe.py:288-295
                _, old_pred_original_sample = self.scheduler.step_cfg_dpm(
                    noise_pred.contiguous(), old_pred_original_sample, t, steps_host[i - 1] if i > 0 else None, latents,
                    self._guidance_scale, generator=generator, out=nxt)
            else:
                self.scheduler.step_cfg(noise_pred.contiguous(), t, latents, self._guidance_scale, out=nxt)
            latents, nxt = nxt, latents
            if callback_on_step_end is not None:
                kw = {"latents": latents, "prompt_embeds": prompt_embeds, "negative_prompt_embeds": negative_prompt_embeds}
                cb = callback_on_step_end(self, i, t, {k: kw[k] for k in callback_on_step_end_tensor_inputs})
                latents = cb.get("latents", latents)
                prompt_embeds = cb.get("prompt_embeds", prompt_embeds)
                negative_prompt_embeds = cb.get("negative_prompt_embeds", negative_prompt_embeds)

        if output_type != "latent":
            video = self.decode_latents(latents)
            video = postprocess_video(video, output_type)
        else:
            video = latents
        if not return_dict:
            return (video,)
        return CogVideoXPipelineOutput(frames=video)


def _device_uint8_ok(video: torch.Tensor) -> bool:
    return (video.is_cuda and video.dtype == torch.bfloat16 and video.dim() == 5 and video.shape[1] == 3
            and (video.shape[3] * video.shape[4]) % 4 == 0)


def postprocess_video(video: torch.Tensor, output_type: str = "np"):
    """D/video_processor.py:89-113: [B,C,F,H,W] in [-1,1] -> per-video [F,H,W,C] in [0,1] (np / pt) or PIL lists.

    Extension: output_type "uint8" returns the np.uint8 array [B,F,H,W,C] that the reference's
    export_to_video would compute from the "np" output (`(frame * 255).astype(np.uint8)`, D/utils/export_utils.py:177-178),
    converted on the device by s2v_video_to_uint8 (bit-exact; a quarter of the device->host bytes of the fp32 "np" path).
    The "pil" path uses the same kernel with numpy_to_pil's rounding."""
    if output_type in ("uint8", "pil") and _device_uint8_ok(video):
        from . import ops
        u8 = ops.video_to_uint8(video.contiguous(), round_half_even=(output_type == "pil")).cpu().numpy()
        if output_type == "uint8":
            return u8
        from PIL import Image
        return [[Image.fromarray(f) for f in vid] for vid in u8]
    if output_type == "uint8":
        raise RuntimeError("postprocess_video(output_type='uint8') needs the decoder's bf16 [B,3,F,H,W] output on a B200 "
                           "(there is no host path); use 'np' for host tensors")
    outs = []
    for b in range(video.shape[0]):
        frames = (video[b].permute(1, 0, 2, 3) / 2 + 0.5).clamp(0, 1)  # [F,C,H,W]
        if output_type == "pt":
            outs.append(frames)
        else:
            arr = frames.cpu().permute(0, 2, 3, 1).float().numpy()
            if output_type == "np":
                outs.append(arr)
            elif output_type == "pil":
                from PIL import Image
                outs.append([Image.fromarray((f * 255).round().astype("uint8")) for f in arr])
            else:
                raise ValueError(f"{output_type} does not exist. Please choose one of ['np', 'pt', 'pil']")
    if output_type == "np":
        import numpy as np
        return np.stack(outs)
    if output_type == "pt":
        return torch.stack(outs)
    return outs


def export_to_video(video_frames, output_video_path: Optional[str] = None, fps: int = 10) -> str:
    """D/utils/export_utils.py:143-186 (S/video_generate.py:74-75): write [F,H,W,3] frames as an mp4.  Accepts what the reference
    accepts (float frames in [0,1] -> `(frame * 255).astype(np.uint8)`, or PIL images) and additionally the uint8 frames of
    postprocess_video(output_type="uint8").  imageio is used when present (the reference's preferred backend), else OpenCV's
    mp4v writer exactly like the reference's _legacy_export_to_video (:116-141)."""
    import tempfile

    import numpy as np

    if output_video_path is None:
        output_video_path = tempfile.NamedTemporaryFile(suffix=".mp4").name
    first = video_frames[0]
    if isinstance(first, np.ndarray):
        frames = [f if f.dtype == np.uint8 else (f * 255).astype(np.uint8) for f in video_frames]
    else:   # PIL images
        frames = [np.array(f) for f in video_frames]
    try:
        import imageio

        imageio.plugins.ffmpeg.get_exe()
        with imageio.get_writer(output_video_path, fps=fps) as writer:
            for f in frames:
                writer.append_data(f)
        return output_video_path
    except (ImportError, AttributeError, RuntimeError):
        pass
    import cv2

    h, w, _ = frames[0].shape
    writer = cv2.VideoWriter(output_video_path, cv2.VideoWriter_fourcc(*"mp4v"), fps=fps, frameSize=(w, h))
    for f in frames:
        writer.write(cv2.cvtColor(np.ascontiguousarray(f), cv2.COLOR_RGB2BGR))
    writer.release()
    return output_video_path
