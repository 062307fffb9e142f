"""CogVideoXDDIMScheduler with the reference's surface (D/schedulers/scheduling_ddim_cogvideox.py:180-402):
`config`, `set_timesteps`, `scale_model_input`, `step`, `timesteps`, `init_noise_sigma`, `order` — plus the fused
`step_cfg` used by the B200 loop (CFG combine + DDIM update in ONE kernel, no host<->device sync).

Schedule tables are float64 / int64 host arithmetic, bit-exact with the reference (torch float64 ops are used because
torch.linspace and np.linspace differ in the last bit).  The per-step coefficients are computed on the HOST from the
integer timestep, so — unlike the reference, which indexes a CPU table with a CUDA timestep and pays two device->host
syncs per step — stepping never synchronises.
"""
from __future__ import annotations

import types
from typing import Optional, Tuple, Union

import numpy as np
import torch

from . import _lib as _lib_mod
from . import ops


class DDIMSchedulerOutput:
    def __init__(self, prev_sample, pred_original_sample=None):
        self.prev_sample = prev_sample
        self.pred_original_sample = pred_original_sample


def _rescale_zero_terminal_snr(alphas_cumprod: torch.Tensor) -> torch.Tensor:
    # scheduling_ddim_cogvideox.py:95-123
    root = alphas_cumprod.sqrt()
    first, last = root[0].clone(), root[-1].clone()
    root = root - last
    root = root * (first / (first - last))
    return root**2


class CogVideoXDDIMScheduler:
    order = 1

    def __init__(self, num_train_timesteps: int = 1000, beta_start: float = 0.00085, beta_end: float = 0.0120,
                 beta_schedule: str = "scaled_linear", trained_betas=None, clip_sample: bool = True,
                 set_alpha_to_one: bool = True, steps_offset: int = 0, prediction_type: str = "epsilon",
                 clip_sample_range: float = 1.0, sample_max_value: float = 1.0, timestep_spacing: str = "leading",
                 rescale_betas_zero_snr: bool = False, snr_shift_scale: float = 3.0):
        self.config = types.SimpleNamespace(
            num_train_timesteps=num_train_timesteps, beta_start=beta_start, beta_end=beta_end, beta_schedule=beta_schedule,
            trained_betas=trained_betas, clip_sample=clip_sample, set_alpha_to_one=set_alpha_to_one, steps_offset=steps_offset,
            prediction_type=prediction_type, clip_sample_range=clip_sample_range, sample_max_value=sample_max_value,
            timestep_spacing=timestep_spacing, rescale_betas_zero_snr=rescale_betas_zero_snr, snr_shift_scale=snr_shift_scale)
        if trained_betas is not None:
            self.betas = torch.tensor(trained_betas, dtype=torch.float32)
        elif beta_schedule == "linear":
            self.betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
        elif beta_schedule == "scaled_linear":
            self.betas = torch.linspace(beta_start**0.5, beta_end**0.5, num_train_timesteps, dtype=torch.float64) ** 2
        else:
            raise NotImplementedError(f"{beta_schedule} is not implemented for {self.__class__}")
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.alphas_cumprod = self.alphas_cumprod / (snr_shift_scale + (1 - snr_shift_scale) * self.alphas_cumprod)
        if rescale_betas_zero_snr:
            self.alphas_cumprod = _rescale_zero_terminal_snr(self.alphas_cumprod)
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.init_noise_sigma = 1.0
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].copy().astype(np.int64))
        # How `fp64 scalar * bf16 tensor` is evaluated: torch on CUDA multiplies with the scalar in fp32 ("cuda", what the
        # reference does on a GPU); torch on CPU first casts the scalar to bf16 ("cpu", what the CPU oracle is pinned to).
        self.scalar_semantics = "cuda"

    @classmethod
    def for_cogvideox(cls, snr_shift_scale: float = 1.0):
        """The scheduler config CogVideoX checkpoints ship (DX/scripts/convert_cogvideox_to_diffusers.py:252-265);
        snr_shift_scale 1.0 for 5B, 3.0 for 2B."""
        return cls(snr_shift_scale=snr_shift_scale, beta_end=0.012, beta_schedule="scaled_linear", beta_start=0.00085,
                   clip_sample=False, num_train_timesteps=1000, prediction_type="v_prediction", rescale_betas_zero_snr=True,
                   set_alpha_to_one=True, timestep_spacing="trailing")

    @classmethod
    def from_reference(cls, scheduler):
        """Build from an already constructed reference `CogVideoXDDIMScheduler` (duck-typed: reads its `.config`), keeping
        any `set_timesteps` state — the one-line swap shown in INTEGRATION.md."""
        cfg = scheduler.config
        get = (lambda k, d=None: cfg[k] if k in cfg else d) if isinstance(cfg, dict) else (lambda k, d=None: getattr(cfg, k, d))
        keys = ("num_train_timesteps", "beta_start", "beta_end", "beta_schedule", "trained_betas", "clip_sample", "set_alpha_to_one",
                "steps_offset", "prediction_type", "clip_sample_range", "sample_max_value", "timestep_spacing",
                "rescale_betas_zero_snr", "snr_shift_scale")
        new = cls(**{k: get(k) for k in keys if get(k, None) is not None or k == "trained_betas"})
        n = getattr(scheduler, "num_inference_steps", None)
        if n is not None:
            new.set_timesteps(n, device=getattr(getattr(scheduler, "timesteps", None), "device", None))
        return new

    # ------------------------------------------------------------------ reference surface
    def scale_model_input(self, sample: torch.Tensor, timestep=None) -> torch.Tensor:
        return sample

    def set_timesteps(self, num_inference_steps: int, device: Union[str, torch.device, None] = None):
        n_train = self.config.num_train_timesteps
        if num_inference_steps > n_train:
            raise ValueError(
                f"`num_inference_steps`: {num_inference_steps} cannot be larger than `self.config.train_timesteps`:"
                f" {n_train} as the unet model trained with this scheduler can only handle maximal {n_train} timesteps.")
        self.num_inference_steps = num_inference_steps
        spacing = self.config.timestep_spacing
        if spacing == "linspace":
            ts = np.linspace(0, n_train - 1, num_inference_steps).round()[::-1].copy().astype(np.int64)
        elif spacing == "leading":
            ts = (np.arange(0, num_inference_steps) * (n_train // num_inference_steps)).round()[::-1].copy().astype(np.int64)
            ts += self.config.steps_offset
        elif spacing == "trailing":
            ts = np.round(np.arange(n_train, 0, -(n_train / num_inference_steps))).astype(np.int64)
            ts -= 1
        else:
            raise ValueError(f"{spacing} is not supported. Please make sure to choose one of 'leading' or 'trailing'.")
        self._timesteps_host = [int(t) for t in ts]
        self.timesteps = torch.from_numpy(ts).to(device)

    def coefficients(self, timestep: int) -> Tuple[float, float, float, float]:
        """(sqrt(a_t), sqrt(1-a_t), a, b) of the v-prediction update, fp64 arithmetic as scheduling_ddim_cogvideox.py:365-392,
        returned as the fp32 values the device multiplies with (see `scalar_semantics`)."""
        if self.num_inference_steps is None:
            raise ValueError("Number of inference steps is 'None', you need to run 'set_timesteps' after creating the scheduler")
        if self.config.prediction_type != "v_prediction":
            raise NotImplementedError("the B200 path implements CogVideoX's v_prediction DDIM update only")
        prev = timestep - self.config.num_train_timesteps // self.num_inference_steps
        a_t = self.alphas_cumprod[timestep]
        a_prev = self.alphas_cumprod[prev] if prev >= 0 else self.final_alpha_cumprod
        beta_t = 1 - a_t
        a_coef = ((1 - a_prev) / (1 - a_t)) ** 0.5
        b_coef = a_prev**0.5 - a_t**0.5 * a_coef
        sa, sb = a_t**0.5, beta_t**0.5

        def on_bf16(c):  # coefficient that multiplies the bf16 sample
            c = torch.as_tensor(c, dtype=torch.float64)
            return float(c.to(torch.bfloat16)) if self.scalar_semantics == "cpu" else float(c.to(torch.float32))

        def on_f32(c):
            return float(torch.as_tensor(c, dtype=torch.float64).to(torch.float32))

        return on_bf16(sa), on_f32(sb), on_bf16(a_coef), on_f32(b_coef)

    @staticmethod
    def _as_int(timestep) -> int:
        # a python int never syncs; a device tensor does (kept for drop-in compatibility with the reference call)
        return int(timestep.item()) if isinstance(timestep, torch.Tensor) else int(timestep)

    def step(self, model_output: torch.Tensor, timestep, sample: torch.Tensor, eta: float = 0.0,
             use_clipped_model_output: bool = False, generator=None, variance_noise: Optional[torch.Tensor] = None,
             return_dict: bool = True):
        """x_t -> x_{t-1}: fp32 `model_output`, bf16 `sample` -> fp32 prev_sample (the reference's dtypes)."""
        if self.num_inference_steps is None:
            raise ValueError("Number of inference steps is 'None', you need to run 'set_timesteps' after creating the scheduler")
        sa, sb, a, b = self.coefficients(self._as_int(timestep))
        mo = model_output.float().contiguous()
        prev = torch.empty_like(mo)
        x0 = torch.empty_like(mo)
        ops.ddim_step(mo, sample.contiguous(), prev, sa, sb, a, b, x0_out=x0)
        if not return_dict:
            return (prev, x0)
        return DDIMSchedulerOutput(prev_sample=prev, pred_original_sample=x0)

    # ------------------------------------------------------------------ fused B200 step
    def step_cfg(self, noise_pred: torch.Tensor, timestep, latents: torch.Tensor, guidance: float,
                 out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """noise_pred [2P,...] bf16 (uncond first) + latents [P,...] bf16 -> next latents bf16; one kernel, bit-exact with
        `.float()` -> CFG -> `step` -> `.to(bf16)` of the reference loop (S/custom_cogvideox_pipe.py:266-296)."""
        sa, sb, a, b = self.coefficients(self._as_int(timestep))
        out = torch.empty_like(latents) if out is None else out
        return ops.cfg_ddim_step(noise_pred, latents, out, float(guidance), sa, sb, a, b)


class CogVideoXDPMScheduler(CogVideoXDDIMScheduler):
    """DPM-Solver++ (SDE, second-order multistep) variant the pipeline also accepts (D/schedulers/scheduling_dpm_cogvideox.py;
    loop branch S/custom_cogvideox_pipe.py:288-295).  Same beta / SNR-shift / zero-terminal-SNR tables and timestep spacing as
    the DDIM scheduler; the update draws Gaussian noise (stochastic), with the reference's draw protocol so that a given
    `torch.Generator` yields the same latents: one bf16 `randn` per step, a second one on second-order steps (the second is
    the one that is used, :421-433)."""

    def dpm_coefficients(self, timestep: int, timestep_back: Optional[int]):
        """(sa, sb, m0, m1, m2, m3, m_noise, second_order) — get_variables / get_mult (:306-328) in fp64 0-dim tensors exactly as
        the reference evaluates them, returned as the fp32 values the device multiplies with."""
        if self.num_inference_steps is None:
            raise ValueError("Number of inference steps is 'None', you need to run 'set_timesteps' after creating the scheduler")
        if self.config.prediction_type != "v_prediction":
            raise NotImplementedError("the B200 path implements CogVideoX's v_prediction update only")
        prev = timestep - self.config.num_train_timesteps // self.num_inference_steps
        a_t = self.alphas_cumprod[timestep]
        a_prev = self.alphas_cumprod[prev] if prev >= 0 else self.final_alpha_cumprod
        a_back = self.alphas_cumprod[timestep_back] if timestep_back is not None else None
        lamb = ((a_t / (1 - a_t)) ** 0.5).log()
        lamb_next = ((a_prev / (1 - a_prev)) ** 0.5).log()
        h = lamb_next - lamb
        mult1 = ((1 - a_prev) / (1 - a_t)) ** 0.5 * (-h).exp()
        mult2 = (-2 * h).expm1() * a_prev**0.5
        m2 = m3 = None
        if a_back is not None:
            lamb_previous = ((a_back / (1 - a_back)) ** 0.5).log()
            r = (lamb - lamb_previous) / h
            m2, m3 = 1 + 1 / (2 * r), 1 / (2 * r)
        mult_noise = (1 - a_prev) ** 0.5 * (1 - (-2 * h).exp()) ** 0.5

        def on_bf16(c):   # multiplies a bf16 tensor
            c = torch.as_tensor(c, dtype=torch.float64)
            return float(c.to(torch.bfloat16)) if self.scalar_semantics == "cpu" else float(c.to(torch.float32))

        def on_f32(c):
            return float(torch.as_tensor(c, dtype=torch.float64).to(torch.float32))

        second = a_back is not None and prev >= 0
        return (on_bf16(a_t**0.5), on_f32((1 - a_t) ** 0.5), on_bf16(mult1), on_f32(mult2), on_f32(m2) if second else 1.0,
                on_f32(m3) if second else 0.0, on_bf16(mult_noise), second)

    @staticmethod
    def _draw(sample: torch.Tensor, generator, variance_noise=None) -> torch.Tensor:
        """randn_tensor(sample.shape, generator, device, dtype) (D/utils/torch_utils.py:38-83): drawn on the generator's device."""
        if variance_noise is not None:
            return variance_noise.to(device=sample.device, dtype=sample.dtype).contiguous()
        if isinstance(generator, (list, tuple)):   # one generator per batch element, like randn_tensor (:66-76)
            if len(generator) == 1:
                generator = generator[0]
            elif len(generator) != sample.shape[0]:
                raise ValueError(f"You have passed a list of generators of length {len(generator)}, but requested an effective batch"
                                 f" size of {sample.shape[0]}. Make sure the batch size matches the length of the generators.")
            else:
                rows = [torch.randn((1,) + tuple(sample.shape[1:]), generator=g, device=g.device, dtype=sample.dtype) for g in generator]
                return torch.cat(rows, dim=0).to(sample.device).contiguous()
        gdev = generator.device if generator is not None else sample.device
        return torch.randn(sample.shape, generator=generator, device=gdev, dtype=sample.dtype).to(sample.device).contiguous()

    def _run(self, model_output, cfg_input, old_x0, timestep, timestep_back, sample, guidance, generator, variance_noise, prev_out):
        t = self._as_int(timestep)
        tb = None if timestep_back is None else self._as_int(timestep_back)
        sa, sb, m0, m1, m2, m3, mn, second = self.dpm_coefficients(t, tb)
        sample = sample.contiguous()
        noise = self._draw(sample, generator, variance_noise)
        use_old = second and old_x0 is not None
        if use_old:
            noise = self._draw(sample, generator, None)   # the reference draws again and uses the second draw (:432)
        x0 = torch.empty(sample.shape, device=sample.device, dtype=torch.float32)
        lib = _lib_mod.load()
        _lib_mod.check(lib.s2v_dpm_step(model_output.data_ptr(), int(cfg_input), sample.data_ptr(),
                                        old_x0.contiguous().data_ptr() if use_old else None, noise.data_ptr(), prev_out.data_ptr(),
                                        x0.data_ptr(), sample.numel(), float(guidance), sa, sb, m0, m1, m2, m3, mn,
                                        torch.cuda.current_stream().cuda_stream), "s2v_dpm_step")
        return prev_out, x0

    def step(self, model_output: torch.Tensor, old_pred_original_sample: Optional[torch.Tensor], timestep, timestep_back, sample: torch.Tensor,
             eta: float = 0.0, use_clipped_model_output: bool = False, generator=None, variance_noise: Optional[torch.Tensor] = None,
             return_dict: bool = False):
        """Reference signature (:330-341): fp32 model_output, bf16 sample -> (fp32 prev_sample, fp32 pred_original_sample)."""
        if sample.dtype != torch.bfloat16 or not sample.is_cuda:
            raise RuntimeError("CogVideoXDPMScheduler.step: expected a CUDA bfloat16 sample")
        mo = model_output.float().contiguous()
        prev = torch.empty_like(mo)
        prev, x0 = self._run(mo, 0, old_pred_original_sample, timestep, timestep_back, sample, 1.0, generator, variance_noise, prev)
        if not return_dict:
            return (prev, x0)
        return DDIMSchedulerOutput(prev_sample=prev, pred_original_sample=x0)

    def step_cfg_dpm(self, noise_pred: torch.Tensor, old_pred_original_sample, timestep, timestep_back, latents: torch.Tensor, guidance: float,
                     generator=None, out: Optional[torch.Tensor] = None):
        """Fused `.float()` -> CFG -> step -> `.to(bf16)`: noise_pred [2P, ...] bf16 (uncond first) -> (next latents bf16, x0 fp32)."""
        out = torch.empty_like(latents) if out is None else out
        return self._run(noise_pred.contiguous(), 1, old_pred_original_sample, timestep, timestep_back, latents, guidance, generator, None, out)
