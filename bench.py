"""Headline benchmark: denoising steps / second of the CogVideoX subject-to-video loop on 49-frame 480x720 latents.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg3|cfg2|cfg4|tiny] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

One "step" = one guided denoising update of ONE video = 2 transformer forwards (uncond + cond) + CFG + DDIM.
Every leg goes through the public call a user makes — `CustomCogVideoXPipeline.__call__` (S/custom_cogvideox_pipe.py:125-326):

  value         weak scaling, one prompt per GPU (prompt sharding: no per-step collective, ONE all-gather of the final latents
                inside the timed region).  One pipeline call of K steps with the inputs already resident in HBM.
  e2e           the same metric with HOST (pinned) buffers: one pipeline call PER STEP — host->device copies of that step's
                latents / prompt embeddings / reference latents and a device->host read of the new latents every step.
  roofline      attention kernel: algorithmic FLOPs / average launch duration, CUDA events on the launching stream, taken in a
                second pass of the same K steps (same power regime; the `value` pass itself carries no per-kernel events).
  cfg_sharded   (N even) N/2 prompts, each on a rank PAIR that runs one CFG half per GPU and exchanges the model output once per
                step (NCCL all_gather_into_tensor inside the pair, on a side stream): strong scaling of one video's step.
  cfg4          (N = 4, workload cfg3) BASELINE configs[3]: 720x1280 (S = 50 626), 2 prompts x 2 CFG halves on 4 GPUs.
  video_e2e     loop + tiled 3-D VAE decode + device uint8 conversion + ONE all-gather of the decoded frames, seconds per video
                (BASELINE configs[4] when N = 8: 8 prompts on 8 GPUs).
  gpu_library_baseline   (N = 1) BASELINE.md §4: the reference's arithmetic for one CogVideoXBlock at the workload shape executed
                by stock torch-CUDA library kernels (cuBLASLt + cuDNN/flash SDPA + eager elementwise; the oracle port run on the
                GPU), extrapolated x layers — the existing Blackwell path the hand-written kernels must beat.  NOT the product path.
  cpu_baseline  (N = 1) the same block on the host cores (the oracle port), extrapolated — see --impl reference.

Prints ONE JSON line on rank 0.

--impl reference times the reference algorithm's CPU implementation (the oracle port — the reference is pure Python on torch, so
the "port" executes the same torch-CPU ops) on the host cores of the box: a bounded sample per step (ONE CogVideoXBlock forward
of ONE sequence at the full workload shape), extrapolated to a full step.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[2]: CogVideoX-5B + LoRA (r=128, alpha=64), 49 frames 480x720, CFG batch 2, one prompt per GPU
    "cfg3": dict(model="CogVideoX-5B", heads=48, layers=42, rotary=True, lora=(128, 64.0), frames=49, height=480, width=720, snr=1.0),
    # BASELINE.json configs[1]: CogVideoX-2B, 13 latent frames (49 px frames) 480x720, no LoRA
    "cfg2": dict(model="CogVideoX-2B", heads=30, layers=30, rotary=False, lora=None, frames=49, height=480, width=720, snr=3.0),
    # BASELINE.json configs[3] geometry: CogVideoX-5B + LoRA, 49 frames 720x1280 (S = 50 626), CFG batch 2
    "cfg4": dict(model="CogVideoX-5B", heads=48, layers=42, rotary=True, lora=(128, 64.0), frames=49, height=720, width=1280, snr=1.0),
    # quick functional check (not a bench line)
    "tiny": dict(model="tiny", heads=2, layers=2, rotary=True, lora=(8, 4.0), frames=9, height=64, width=96, snr=1.0),
}
TEXT_LEN, TEXT_DIM = 226, 4096
GUIDANCE = 6.0


def geometry(w):
    n = (w["height"] // 16) * (w["width"] // 16)
    F = (w["frames"] - 1) // 4 + 1
    S = TEXT_LEN + n + F * n
    D = w["heads"] * 64
    return n, F, S, D


def flops_per_step(w):
    """Algorithmic FLOPs of one guided step: per sample-layer 24 S D^2 + 4 S^2 D (+ 36 S D r)."""
    n, F, S, D = geometry(w)
    r = w["lora"][0] if w["lora"] else 0
    per = 24 * S * D * D + 4 * S * S * D + 36 * S * D * r
    return 2 * w["layers"] * per, 2 * w["layers"] * 4 * S * S * D


def config_of(args, w, prompts):
    """The `config` object both arms print (same workload name for the product and the reference arm)."""
    n, F, S, D = geometry(w)
    total_fl, _ = flops_per_step(w)
    return {"workload": f"{args.workload}: {w['model']}{' + LoRA r=%d' % w['lora'][0] if w['lora'] else ''}, {w['frames']} frames "
                        f"{w['height']}x{w['width']}, CFG batch 2, {w['layers']} layers, S={S} tokens, one prompt per GPU",
            "prompts": prompts, "sharding": "prompt (no per-step collective; one all-gather of final latents)",
            "l2": "inputs larger than L2: 11+ GB weights and 235 MB activations per tensor vs 126 MB L2",
            "tflop_per_step": round(total_fl / 1e12, 1)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", 1430.0), d.get("hbm_gbs", 6459.0), "measured (MEASURED_PEAKS.json, sustained)"
    return 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(workload):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the roofline kernel, read from the committed summary of an
    `ncu --set full` capture at this shape (profiles/ncu_traffic.json, written by tools/ncu_summary.py --traffic)."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(p):
        return None, None
    d = json.load(open(p)).get("attn_fwd_kernel", {}).get(workload)
    return (d["dram_bytes"], d["source"]) if d else (None, None)


class EnergyMeter:
    """Board energy over an interval from NVML's cumulative counter (nvmlDeviceGetTotalEnergyConsumption, millijoules)."""

    def __init__(self, index):
        self.h = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.read()
        except Exception:
            self.h = None

    def read(self):
        return self.nv.nvmlDeviceGetTotalEnergyConsumption(self.h) / 1e3 if self.h is not None else None


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        num = lambda s: s.replace(".", "").isdigit()  # noqa: E731
        sm = [float(r[1]) for r in self.rows if len(r) >= 8 and num(r[1])]
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and num(r[2])]
        pw = [float(r[3]) for r in self.rows if len(r) >= 8 and num(r[3])]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 8 for i in range(4) if r[4 + i].lower().startswith("active")})
        busy = [s for s in sm if s > 0.5 * (max(sm) if sm else 1)]
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm), "power_w_median": statistics.median(pw) if pw else None}


# ---------------------------------------------------------------------------------------------- product arm
def build_model(w, dev):
    import torch

    import s2v_b200
    t0 = time.time()
    kw = dict(num_attention_heads=w["heads"], num_layers=w["layers"], use_rotary_positional_embeddings=w["rotary"])
    if w["model"] == "tiny":
        kw.update(time_embed_dim=64, text_embed_dim=64)
    with torch.device("meta"):
        m = s2v_b200.CogVideoXTransformer3DModel(**kw).to(torch.bfloat16)
        if w["lora"]:
            s2v_b200.inject_lora(m, w["lora"][0], w["lora"][1])
    m = m.to_empty(device=dev)
    g = torch.Generator(device=dev).manual_seed(0)
    with torch.no_grad():
        for name, p in m.named_parameters():
            ln = (".norm." in name) or name.startswith("norm_final") or "norm_q" in name or "norm_k" in name
            if ln and name.endswith("weight"):
                p.copy_(1.0 + 0.1 * torch.randn(p.shape, device=dev, generator=g))
            elif ln:
                p.copy_(0.1 * torch.randn(p.shape, device=dev, generator=g))
            else:
                p.copy_(0.02 * torch.randn(p.shape, device=dev, generator=g))
    m.invalidate_engine()
    m.engine()
    torch.cuda.synchronize()
    return m, time.time() - t0


def build_vae(dev):
    import torch

    import s2v_b200
    with torch.device("meta"):
        vae = s2v_b200.AutoencoderKLCogVideoX(scaling_factor=0.7)
    vae = vae.to_empty(device=dev)
    g = torch.Generator(device=dev).manual_seed(3)
    with torch.no_grad():
        for n, p in vae.named_parameters():
            if "norm_layer.weight" in n:
                p.copy_(1 + 0.1 * torch.randn(p.shape, device=dev, generator=g))
            elif n.endswith("bias"):
                p.copy_(0.05 * torch.randn(p.shape, device=dev, generator=g))
            else:
                p.copy_(torch.randn(p.shape, device=dev, generator=g) / p[0].numel() ** 0.5)
    vae = vae.to(torch.bfloat16)
    vae.enable_slicing()
    vae.enable_tiling()
    return vae


class Inputs:
    """Synthetic inputs of one rank in PINNED host memory (SURVEY §8d shapes / scales) and their device copies."""

    def __init__(self, w, P, dev, seed, text_dim):
        import torch
        n, F, S, D = geometry(w)
        h8, w8 = w["height"] // 8, w["width"] // 8
        g = torch.Generator().manual_seed(seed)
        bf = torch.bfloat16
        self.lat_h = torch.randn(P, F, 16, h8, w8, generator=g).to(bf).pin_memory()
        self.neg_h = (0.2 * torch.randn(P, TEXT_LEN, text_dim, generator=g)).to(bf).pin_memory()
        self.pos_h = (0.2 * torch.randn(P, TEXT_LEN, text_dim, generator=g)).to(bf).pin_memory()
        self.ref_h = (0.7 * torch.randn(P, 1, 16, h8, w8, generator=g)).to(bf).pin_memory()
        self.out_h = torch.empty_like(self.lat_h).pin_memory()
        self.lat_d, self.neg_d, self.pos_d, self.ref_d = (t.to(dev) for t in (self.lat_h, self.neg_h, self.pos_h, self.ref_h))
        self.h2d_bytes = int(self.lat_h.nbytes + self.neg_h.nbytes + self.pos_h.nbytes + self.ref_h.nbytes)
        self.d2h_bytes = int(self.out_h.nbytes)


def run_product(args, w):
    import torch
    import torch.distributed as dist

    import s2v_b200
    from s2v_b200 import _lib, ops, parallel

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {args.gpus}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.check(_lib.load().s2v_device_check(local), "s2v_device_check")

    text_dim = 64 if w["model"] == "tiny" else TEXT_DIM
    model, t_build = build_model(w, dev)
    sched = s2v_b200.CogVideoXDDIMScheduler.for_cogvideox(w["snr"])
    sched.set_timesteps(50)
    all_steps = list(sched._timesteps_host)
    pipe = s2v_b200.CustomCogVideoXPipeline(None, None, model, None, sched)

    def window(first, count):
        return [all_steps[(first + i) % len(all_steps)] for i in range(count)]

    def call(wl, lat, pos, neg, ref, ts, **kw):
        """The public call (S/video_generate.py:47-61 form), latents out."""
        return pipe(prompt=None, prompt_embeds=pos, negative_prompt_embeds=neg, ref_img_states=ref, latents=lat, height=wl["height"],
                    width=wl["width"], num_frames=wl["frames"], num_inference_steps=50, timesteps=ts, guidance_scale=GUIDANCE,
                    output_type="latent", return_dict=False, eval=True, **kw)[0]

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    meter = EnergyMeter(local) if rank == 0 else None
    energy = {}

    def timed_loop(wl, inp, sp, P_total, steps, warmup, host_io=False, timer=None, energy_key=None):
        """W untimed + K timed guided steps of this rank's prompts, then the path's single collective; device time, max over ranks."""
        lat = call(wl, inp.lat_d, inp.pos_d, inp.neg_d, inp.ref_d, window(0, warmup)) if warmup else inp.lat_d
        sync_all()
        ops.set_kernel_timer(timer)
        launches0 = _lib.launch_count
        j0 = meter.read() if (meter and energy_key) else None
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.nvtx.range_push("timed")   # `ncu --nvtx --nvtx-include timed/` captures exactly the timed launches
        e0.record()
        if host_io:
            # one public call per step with HOST buffers: inputs copied in, the new latents copied out, every step
            for i in range(steps):
                lat = call(wl, inp.lat_h, inp.pos_h, inp.neg_h, inp.ref_h, window(warmup + i, 1))
                inp.out_h.copy_(lat, non_blocking=True)
                torch.cuda.current_stream().synchronize()   # the caller owns the result on the host before the next step
        else:
            lat = call(wl, lat, inp.pos_d, inp.neg_d, inp.ref_d, window(warmup, steps))
        gathered = parallel.gather_latents(lat, P_total, sp) if world > 1 else lat   # the path's single collective
        e1.record()
        torch.cuda.nvtx.range_pop()
        sync_all()
        if j0 is not None:
            energy[energy_key] = (meter.read() - j0) / steps
        ops.set_kernel_timer(None)
        ms = max_over_ranks(e0.elapsed_time(e1))
        assert torch.isfinite(gathered.float()).all(), "non-finite latents"
        return ms, _lib.launch_count - launches0, lat

    # ------------------------------------------------------------ headline: weak scaling, one prompt per GPU
    P_total = world
    sp = parallel.plan(P_total, world, rank, mode="prompt")
    P = len(sp.prompts)
    inp = Inputs(w, P, dev, 100 + rank, text_dim)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms, launches, lat_final = timed_loop(w, inp, sp, P_total, args.steps, args.warmup, energy_key="value_loop")
    clocks = sampler.stop() if rank == 0 else None
    # second pass of the same steps with CUDA events around the five big kernels (rank 0 only records)
    names = ["s2v_attn_fwd", "s2v_qkv_lora", "s2v_outproj_lora_gate_residual", "s2v_ffn_up_gelu_lora", "s2v_ffn_down_lora_gate_residual"]
    timer = ops.KernelTimer(names) if rank == 0 else None
    ms_prof, _, _ = timed_loop(w, inp, sp, P_total, min(args.steps, 5), 1, timer=timer)
    kern = timer.summary() if timer else {}
    prof_steps = min(args.steps, 5)
    ms_e2e = timed_loop(w, inp, sp, P_total, args.steps, 1, host_io=True)[0] if args.e2e else float("nan")   # --no-e2e: profiler runs only

    line = None
    if rank == 0:
        n, F, S, D = geometry(w)
        total_fl, attn_fl = flops_per_step(w)
        peak_tf, peak_gb, peak_src = measured_peaks()
        value = world * args.steps / (ms / 1e3)
        e2e_v = world * args.steps / (ms_e2e / 1e3)
        attn = kern.get("s2v_attn_fwd")
        attn_per_launch = attn_fl / (2 * w["layers"]) * (2 * P)      # 4 S^2 D per sequence, 2P sequences per launch
        roof = None
        if attn:
            ach = attn_per_launch / (attn["avg_ms"] / 1e3) / 1e12
            traffic, traffic_src = ncu_traffic(args.workload)
            roof = {"kernel": "attn_fwd_kernel", "bound": "tensor", "achieved": round(ach, 1), "peak": peak_tf, "unit": "TFLOP/s",
                    "frac": round(ach / peak_tf, 4), "traffic": traffic, "traffic_source": traffic_src,
                    "algorithmic_bytes": 3 * S * D * 2 * 2 * P + S * D * 2 * 2 * P, "peak_source": peak_src,
                    "avg_launch_ms": round(attn["avg_ms"], 4), "launches_timed": attn["launches"],
                    "share_of_step": round(attn["total_ms"] / ms_prof, 4),
                    "measured_in": f"a second pass of {prof_steps} steps of the same loop with CUDA events around the launches "
                                   f"({ms_prof / prof_steps:.1f} ms/step with the events)"}
            # the unit that actually bounds a head_dim-64 softmax: MUFU.EX2 issues 16 results / clk / SM (tools/microbench_mufu_mix.cu)
            # against 8192 MMA flop / clk / SM, so one exponential per 256 MMA flop needs twice the cycles of its MMAs; the kernel
            # sends 7 of 8 exponentials there.  Peak taken at the median SM clock sampled during the timed region.
            sm_mhz = float((clocks or {}).get("sm_mhz") or 0)
            if sm_mhz > 0:
                exps = 2 * P * w["heads"] * float(S) * S * 7 / 8
                xu_peak = 16 * 148 * sm_mhz * 1e6
                roof["xu"] = {"mufu_ex2_per_s": round(exps / (attn["avg_ms"] / 1e3) / 1e12, 3), "peak_per_s": round(xu_peak / 1e12, 3),
                              "unit": "T exp2/s", "frac": round(exps / (attn["avg_ms"] / 1e3) / xu_peak, 3),
                              "tensor_frac_ceiling_if_xu_saturated": round(256 * xu_peak * 8 / 7 / 1e12 / peak_tf, 3)}
        gemm_fl = {"s2v_qkv_lora": 2 * S * D * (3 * D), "s2v_outproj_lora_gate_residual": 2 * S * D * D,
                   "s2v_ffn_up_gelu_lora": 2 * S * D * 4 * D, "s2v_ffn_down_lora_gate_residual": 2 * S * D * 4 * D}
        kernels = {k: {"avg_ms": round(v["avg_ms"], 4), "share_of_step": round(v["total_ms"] / ms_prof, 4),
                       **({"tflops": round(gemm_fl[k] * 2 * P / (v["avg_ms"] / 1e3) / 1e12, 1)} if k in gemm_fl else {})}
                   for k, v in kern.items()}
        line = {
            "metric": "denoising_steps_per_sec", "value": round(value, 4),
            "unit": "steps/s (one step = one guided update of one 49f 480x720 video)",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 3),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": config_of(args, w, P_total),
            "clocks": clocks,
            "energy": {"joule_per_step_gpu0": round(energy["value_loop"], 1) if energy.get("value_loop") else None,
                       "mean_power_w": round(energy["value_loop"] / (ms / args.steps / 1e3), 1) if energy.get("value_loop") else None,
                       "source": "nvmlDeviceGetTotalEnergyConsumption around the timed loop (rank 0's GPU)"},
            "e2e": {"value": round(e2e_v, 4), "unit": "steps/s", "h2d_bytes_per_step": inp.h2d_bytes, "d2h_bytes_per_step": inp.d2h_bytes,
                    "ms_per_step": round(ms_e2e / args.steps, 3),
                    "api": "CustomCogVideoXPipeline.__call__ once per step with pinned host tensors (latents, prompt embeddings, reference latents in; latents out)"},
            "gpu_launches": launches,
            "roofline": roof,
            "kernels": kernels,
            "model_tflops": round(total_fl * world * args.steps / (ms / 1e3) / 1e12 / world, 1),
            "build_s": round(t_build, 1),
        }

    # ------------------------------------------------------------ CFG-sharded sub-runs (strong scaling of one video's step)
    def cfg_sharded_leg(wl, steps, warmup, tag):
        pairs = world // 2
        spc = parallel.plan(pairs, world, rank, mode="cfg")
        inpc = Inputs(wl, len(spc.prompts), dev, 500 + rank // 2, text_dim)      # both ranks of a pair draw the same prompt
        pipe.enable_cfg_parallel(spc, record_times=True)
        try:
            msc, _, _ = timed_loop(wl, inpc, spc, pairs, steps, warmup)
            xt = pipe._cfg_xchg.times_ms()[warmup:]
        finally:
            pipe.disable_cfg_parallel()
        n_, F_, S_, D_ = geometry(wl)
        fl, _ = flops_per_step(wl)
        xbytes = int(inpc.lat_d[:1].numel() * 2 * len(spc.prompts))
        return {"workload": tag, "prompts": pairs, "gpus_per_prompt": 2, "steps": steps, "warmup": warmup,
                "ms_per_step_per_video": round(msc / steps, 3), "steps_per_s_aggregate": round(pairs * steps / (msc / 1e3), 4),
                "scaling": "strong (one prompt's CFG batch of 2 split over 2 GPUs)", "tokens": S_,
                "model_tflops_per_gpu": round(fl * steps / 2 / (msc / 1e3) / 1e12, 1),
                "collective": "NCCL all_gather_into_tensor inside each rank pair, once per step, on a side stream; + one all-gather of "
                              "the final latents",
                "exchange_bytes_per_rank_per_step": xbytes,
                "exchange_ms_per_step": {"median": round(statistics.median(xt), 4), "max": round(max(xt), 4)} if xt else None}

    if world >= 2 and world % 2 == 0 and args.sub_runs:
        r = cfg_sharded_leg(w, min(args.steps, 8), 3, f"{args.workload} geometry, CFG halves on rank pairs")
        if rank == 0:
            line["cfg_sharded"] = r
        if world == 4 and args.workload == "cfg3":
            model.engine()._ws.clear()
            r = cfg_sharded_leg(WORKLOADS["cfg4"], 2, 1, "cfg4: CogVideoX-5B + LoRA r=128, 49 frames 720x1280 (S=50626), 2 prompts x 2 CFG "
                                                        "halves on 4 GPUs (BASELINE configs[3])")
            if rank == 0:
                line["cfg4"] = r
            model.engine()._ws.clear()

    # ------------------------------------------------------------ loop + VAE decode + all-gather of frames: seconds per video
    if args.vae and w["model"] != "tiny" and args.sub_runs:
        r = video_e2e_leg(args, w, pipe, inp, sp, P_total, lat_final, ms / args.steps, dev, world, rank, sync_all, max_over_ranks)
        if rank == 0:
            line["video_e2e"] = r
            line["vae_decode"] = r["vae_decode"]
    if rank == 0:
        if args.library_baseline and world == 1 and w["model"] != "tiny":
            del inp
            line["gpu_library_baseline"] = gpu_library_leg(w, dev, line["ms_per_step"])
        if args.cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_sample(w, steps=1, warmup=0)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def video_e2e_leg(args, w, pipe, inp, sp, P_total, latents, ms_per_step, dev, world, rank, sync_all, max_over_ranks):
    """Row V next to the loop (BASELINE configs[4]): every rank decodes ITS video with the 3-D causal VAE (random-init CogVideoX
    decoder, the reference's default tiled schedule, D/models/autoencoders/autoencoder_kl_cogvideox.py:1374-1455), converts the
    frames to uint8 on the device (export_to_video's rounding) and the ranks all-gather the frames — the path's single collective,
    here carrying decoded frames instead of latents."""
    import torch
    import torch.distributed as dist

    import s2v_b200
    from s2v_b200 import ops, vae as vae_mod

    vae = build_vae(dev)
    vpipe = s2v_b200.CustomCogVideoXPipeline(None, None, None, vae, None)
    P = latents.shape[0]
    times = []
    flops = {}
    for it in range(2):
        sync_all()
        vae_mod.CONV_FLOPS = flops if it == 0 else None
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        video = vpipe.decode_latents(latents)                                  # [P, 3, 49, H, W] bf16
        frames = ops.video_to_uint8(video.contiguous(), round_half_even=False)  # [P, 49, H, W, 3] uint8
        e1.record()
        if world > 1:
            allf = torch.empty((world * P,) + tuple(frames.shape[1:]), dtype=torch.uint8, device=dev)
            dist.all_gather_into_tensor(allf, frames)
        else:
            allf = frames
        e2.record()
        sync_all()
        times.append((max_over_ranks(e0.elapsed_time(e1)), max_over_ranks(e1.elapsed_time(e2))))
    vae_mod.CONV_FLOPS = None
    assert torch.isfinite(video.float()).all() and allf.shape[0] == P_total
    dec_ms, gat_ms = min(t[0] for t in times), min(t[1] for t in times)
    peak_tf, _, _ = measured_peaks()
    conv_tf = flops.get("algorithmic", 0.0) / P / (dec_ms / 1e3) / 1e12
    loop50 = 50 * ms_per_step / 1e3
    return {"videos": P_total, "gpus": world, "denoise_s_per_video_50_steps": round(loop50, 3),
            "decode_plus_uint8_s_per_video": round(dec_ms / 1e3, 4), "allgather_frames_ms": round(gat_ms, 3),
            "allgather_bytes_per_rank": int(frames.numel()),
            "seconds_per_video": round(loop50 + dec_ms / 1e3 + gat_ms / 1e3, 3),
            "videos_per_s_aggregate": round(P_total / (loop50 + dec_ms / 1e3 + gat_ms / 1e3), 4),
            "note": "50 DDIM steps (the reference's default, S/video_generate.py:58) at the measured ms/step of the timed loop + measured "
                    "decode + measured all-gather; T5 prompt encode and reference-image VAE encode are outside (SURVEY §8d)",
            "vae_decode": {"seconds_per_video": round(dec_ms / 1e3 / P, 4), "schedule": "tiled 3x3 x 6 temporal batches (reference default), bf16, "
                           "+ device uint8 conversion", "output": list(video.shape),
                           "conv_flop_algorithmic": flops.get("algorithmic"), "conv_flop_launched_incl_border_ring": flops.get("launched"),
                           "conv_launches": flops.get("launches"),
                           "conv_tflops": round(conv_tf, 1), "frac_of_sustained_bf16_peak": round(conv_tf / peak_tf, 3),
                           "flop_source": "sum over the launched implicit-GEMM convolutions of 2 x T x H x W x Cout x taps x Cin"}}


# ---------------------------------------------------------------------------------------------- torch-CUDA library leg
def gpu_library_leg(w, dev, product_ms_per_step):
    """BASELINE.md §4 "B200 library reference point": the reference's block arithmetic (the oracle port = the same torch ops the
    reference modules issue) executed by torch-CUDA eager on this GPU — cuBLASLt GEMMs, cuDNN / flash SDPA, ATen elementwise and
    the reference's cat / split copies — for ONE CogVideoXBlock at the workload shape with the CFG batch of 2, extrapolated x
    layers.  Also interleaved A/Bs of the two dominant kernels against their library counterparts on identical tensors."""
    import torch
    import torch.nn.functional as F

    from oracle import s2v_oracle as O      # measurement only: the product path never imports the oracle
    from s2v_b200 import ops

    n, Fr, S, D = geometry(w)
    bf = torch.bfloat16
    cfg = O.TransformerConfig(num_attention_heads=w["heads"], num_layers=1, use_rotary_positional_embeddings=w["rotary"],
                              lora_rank=w["lora"][0] if w["lora"] else 0, lora_alpha=w["lora"][1] if w["lora"] else 0.0)
    p = {k: v.to(device=dev, dtype=bf) for k, v in O.synth_params(cfg, seed=0).items()}
    g = torch.Generator().manual_seed(1)
    vid, txt, ref = (torch.randn(2, m, D, generator=g).to(bf).to(dev) for m in (Fr * n, TEXT_LEN, n))
    temb = torch.randn(2, cfg.time_embed_dim, generator=g).to(bf).to(dev)
    rv = rr = None
    if w["rotary"]:
        rv, rr = O.pipeline_rope_tables(w["height"], w["width"], Fr)
        rv, rr = tuple(t.to(dev) for t in rv), tuple(t.to(dev) for t in rr)

    def block():
        return O.block_forward(p, cfg, "transformer_blocks.0.", vid, txt, temb, ref, rv, rr)

    def timed(fn, iters):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters

    with torch.no_grad():
        blk_ms = timed(block, 20)        # ~20 blocks back to back: the sustained (power-capped) regime of a real step
    lib_ms_step = blk_ms * w["layers"]
    out = {"kind": "torch-CUDA eager (cuBLASLt + cuDNN/flash SDPA + ATen elementwise) executing the reference's block arithmetic",
           "sample": f"1 CogVideoXBlock forward, CFG batch 2, S={S} D={D}, bf16, {blk_ms:.2f} ms/block sustained; extrapolated x{w['layers']} layers",
           "ms_per_step": round(lib_ms_step, 1), "steps_per_s": round(1e3 / lib_ms_step, 4),
           "product_speedup_vs_library": round(lib_ms_step / product_ms_per_step, 3)}
    # interleaved kernel A/Bs on identical tensors (alternating order, medians)
    H = w["heads"]
    qkv = torch.randn(2, S, 3 * D, device=dev).to(bf)
    o = torch.empty(2, S, D, device=dev, dtype=bf)
    q4 = qkv.view(2, S, 3, H, 64)
    qs, ks, vs = (q4[:, :, i].transpose(1, 2) for i in range(3))
    x = torch.randn(2 * S, D, device=dev).to(bf)
    wt = (0.02 * torch.randn(4 * D, D, device=dev)).to(bf)
    bias = (0.02 * torch.randn(4 * D, device=dev)).to(bf)
    y = torch.empty(2 * S, 4 * D, device=dev, dtype=bf)
    arms = {"attn_s2v": lambda: ops.attention(qkv, o, H), "attn_sdpa": lambda: F.scaled_dot_product_attention(qs, ks, vs),
            "ffn_up_s2v_linear_bias": lambda: ops.linear(x, wt, bias, y), "ffn_up_torch_addmm": lambda: torch.addmm(bias, x, wt.t(), out=y)}
    res = {k: [] for k in arms}
    order = list(arms)
    for rep in range(6):
        for k in (order if rep % 2 == 0 else order[::-1]):
            res[k].append(timed(arms[k], 6))
    med = {k: statistics.median(v) for k, v in res.items()}
    fl_attn, fl_gemm = 4.0 * S * S * D * 2, 2.0 * 2 * S * D * 4 * D
    out["kernel_ab"] = {"method": "in-process, interleaved, alternating order, 6 repetitions x 6 launches, medians; isolated (not in-step)",
                        "attention": {"s2v_attn_fwd_ms": round(med["attn_s2v"], 3), "torch_sdpa_ms": round(med["attn_sdpa"], 3),
                                      "s2v_tflops": round(fl_attn / med["attn_s2v"] / 1e9, 1), "sdpa_tflops": round(fl_attn / med["attn_sdpa"] / 1e9, 1),
                                      "s2v_over_sdpa_time": round(med["attn_s2v"] / med["attn_sdpa"], 3)},
                        "linear_ffn_up": {"s2v_linear_ms": round(med["ffn_up_s2v_linear_bias"], 3), "torch_addmm_ms": round(med["ffn_up_torch_addmm"], 3),
                                          "s2v_tflops": round(fl_gemm / med["ffn_up_s2v_linear_bias"] / 1e9, 1),
                                          "cublas_tflops": round(fl_gemm / med["ffn_up_torch_addmm"] / 1e9, 1),
                                          "s2v_over_cublas_time": round(med["ffn_up_s2v_linear_bias"] / med["ffn_up_torch_addmm"], 3)}}
    return out


# ---------------------------------------------------------------------------------------------- CPU arm (oracle port)
def cpu_sample(w, steps, warmup):
    """ONE CogVideoXBlock forward of ONE sequence at the workload's full shape on the host cores, bf16 (the reference's
    dtype), all threads; a guided step = 2 sequences x `layers` blocks, so steps/s = 1 / (2 * layers * t_block)."""
    import torch

    from oracle import s2v_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    n, F, S, D = geometry(w)
    cfg = O.TransformerConfig(num_attention_heads=w["heads"], num_layers=1, use_rotary_positional_embeddings=w["rotary"],
                              lora_rank=w["lora"][0] if w["lora"] else 0, lora_alpha=w["lora"][1] if w["lora"] else 0.0,
                              **(dict(time_embed_dim=64, text_embed_dim=64) if w["model"] == "tiny" else {}))
    p = {k: v.to(torch.bfloat16) for k, v in O.synth_params(cfg, seed=0).items()}
    g = torch.Generator().manual_seed(1)
    bf = torch.bfloat16
    vid, txt, ref = (torch.randn(1, m, D, generator=g).to(bf) for m in (F * n, TEXT_LEN, n))
    temb = torch.randn(1, cfg.time_embed_dim, generator=g).to(bf)
    rv = rr = None
    if w["rotary"]:
        rv, rr = O.pipeline_rope_tables(w["height"], w["width"], F)
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.time()
            O.block_forward(p, cfg, "transformer_blocks.0.", vid, txt, temb, ref, rv, rr)
            if i >= warmup:
                times.append(time.time() - t0)
    t_block = sum(times) / len(times)
    v = 1.0 / (2 * w["layers"] * t_block)
    return {"value": v, "unit": "steps/s", "cores": cores, "kind": "port",
            "sample": f"1 CogVideoXBlock forward, 1 of 2 CFG sequences, full shape S={S} D={D}, torch-CPU bf16, {t_block:.2f} s/block; "
                      f"extrapolated x{2 * w['layers']} to a guided step", "block_s": t_block}


def run_reference(args, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n, F, S, D = geometry(w)
    res = cpu_sample(w, steps=max(1, min(args.steps, 8)), warmup=min(args.warmup, 1))
    total_fl, _ = flops_per_step(w)
    line = {"impl": "reference", "metric": "denoising_steps_per_sec", "value": res["value"],
            "unit": "steps/s (one step = one guided update of one 49f 480x720 video)", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 / res["value"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": config_of(args, w, args.gpus),
            "cpu_baseline": res, "e2e": {"value": res["value"], "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", dest="cpu_baseline", action="store_false")
    ap.add_argument("--no-library-baseline", dest="library_baseline", action="store_false",
                    help="skip the torch-CUDA library leg (gpu_library_baseline; N=1 only)")
    ap.add_argument("--no-vae", dest="vae", action="store_false", help="skip the loop + VAE decode + all-gather leg (video_e2e)")
    ap.add_argument("--no-sub-runs", dest="sub_runs", action="store_false", help="headline only (no cfg_sharded / cfg4 / video_e2e)")
    ap.add_argument("--no-e2e", dest="e2e", action="store_false", help="skip the host-buffer loop (profiler runs; not a valid bench line)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, w)
    else:
        run_product(args, w)
