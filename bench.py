"""Headline benchmark: denoising steps / second of the CogVideoX subject-to-video loop on 49-frame 480x720 latents.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg3|cfg2|tiny] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

One "step" = one guided denoising update of ONE video = 2 transformer forwards (uncond + cond) + CFG + DDIM.
Weak scaling: every GPU owns one prompt (prompt sharding, no per-step collective; one all-gather of the final
latents inside the timed region).  `value` = (videos x steps) / device time with inputs resident in HBM; `e2e` = the
same through the public pipeline objects with HOST (pinned) buffers copied in and the new latents copied out every
step.  Prints ONE JSON line on rank 0.

--impl reference times the reference algorithm's CPU implementation (the oracle port — the reference is pure Python on
torch, so the "port" executes the same torch-CPU ops) on the host cores of the box: a bounded sample per step (ONE
CogVideoXBlock forward of ONE sequence at the full workload shape), extrapolated to a full step.
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
    # BASELINE.json configs[3] geometry on one GPU: CogVideoX-5B + LoRA, 49 frames 720x1280 (S = 50 626), CFG batch 2
    "cfg4": dict(model="CogVideoX-5B", heads=48, layers=42, rotary=True, lora=(128, 64.0), frames=49, height=720, width=1280, snr=1.0),
    # quick functional check (not a bench line)
    "tiny": dict(model="tiny", heads=2, layers=2, rotary=True, lora=(8, 4.0), frames=9, height=64, width=96, snr=1.0),
}
TEXT_LEN, TEXT_DIM = 226, 4096


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
        sm = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 8 for i in range(4) if r[4 + i].lower().startswith("active")})
        busy = [s for s in sm if s > 0.5 * (max(sm) if sm else 1)]
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


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
    m.engine()
    torch.cuda.synchronize()
    return m, time.time() - t0


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

    n, F, S, D = geometry(w)
    text_dim = 64 if w["model"] == "tiny" else TEXT_DIM
    model, t_build = build_model(w, dev)
    sched = s2v_b200.CogVideoXDDIMScheduler.for_cogvideox(w["snr"])
    sched.set_timesteps(50)
    pipe = s2v_b200.CustomCogVideoXPipeline(None, None, model, None, sched)
    P_total = world                      # weak scaling: one prompt per GPU
    sp = parallel.plan(P_total, world, rank)
    P = len(sp.prompts)
    h8, w8 = w["height"] // 8, w["width"] // 8
    g = torch.Generator().manual_seed(100 + rank)
    # synthetic inputs in PINNED host memory (SURVEY §8d shapes / scales)
    lat_h = torch.randn(P, F, 16, h8, w8, generator=g).to(torch.bfloat16).pin_memory()
    pe_h = (0.2 * torch.randn(2 * P, TEXT_LEN, text_dim, generator=g)).to(torch.bfloat16).pin_memory()
    ref_h = (0.7 * torch.randn(P, 1, 16, h8, w8, generator=g)).to(torch.bfloat16).pin_memory()
    out_h = torch.empty_like(lat_h).pin_memory()
    rope = pipe.rotary_tables(w["height"], w["width"], F, dev) if w["rotary"] else None
    img_rope = (rope[0][n:], rope[1][n:]) if rope else None
    ref_rope = (rope[0][:n], rope[1][:n]) if rope else None
    steps_host = sched._timesteps_host
    t_dev = torch.tensor(steps_host, device=dev, dtype=torch.float32)

    lat_d, pe_d, ref_d = lat_h.to(dev), pe_h.to(dev), ref_h.to(dev)
    model_in = torch.empty((2 * P,) + tuple(lat_d.shape[1:]), device=dev, dtype=torch.bfloat16)
    nxt = torch.empty_like(lat_d)

    def one_step(i, lat, host_io):
        nonlocal nxt
        if host_io:  # e2e: this step's inputs come from host memory, the result goes back to host memory
            lat.copy_(lat_h, non_blocking=True)
            pe_d.copy_(pe_h, non_blocking=True)
            ref_d.copy_(ref_h, non_blocking=True)
        model_in[:P].copy_(lat)
        model_in[P:].copy_(lat)
        k = i % len(steps_host)
        noise = model(hidden_states=model_in, encoder_hidden_states=pe_d, ref_img_states=ref_d, timestep=t_dev[k:k + 1].expand(2 * P),
                      image_rotary_emb=img_rope, ref_image_rotary_emb=ref_rope, return_dict=False, eval=True)[0]
        sched.step_cfg(noise, steps_host[k], lat, 6.0, out=nxt)
        if host_io:
            out_h.copy_(nxt, non_blocking=True)
            torch.cuda.current_stream().synchronize()   # the caller owns the result on the host before the next step
        return nxt

    def timed_loop(host_io, timer=None):
        nonlocal nxt, lat_d
        for i in range(args.warmup):
            new = one_step(i, lat_d, host_io)
            nxt, lat_d = lat_d, new
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ops.set_kernel_timer(timer)
        launches0 = _lib.launch_count
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.nvtx.range_push("timed")   # `ncu --nvtx --nvtx-include timed/` captures exactly the timed launches
        e0.record()
        for i in range(args.steps):
            new = one_step(args.warmup + i, lat_d, host_io)
            nxt, lat_d = lat_d, new
        gathered = parallel.gather_latents(lat_d, P_total, sp) if world > 1 else lat_d   # the path's single collective
        e1.record()
        torch.cuda.nvtx.range_pop()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ops.set_kernel_timer(None)
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        assert torch.isfinite(gathered.float()).all(), "non-finite latents"
        return ms, _lib.launch_count - launches0

    sampler = ClockSampler(local)
    names = ["s2v_attn_fwd", "s2v_qkv_lora", "s2v_outproj_lora_gate_residual", "s2v_ffn_up_gelu_lora", "s2v_ffn_down_lora_gate_residual"]
    timer = ops.KernelTimer(names) if rank == 0 else None
    if rank == 0:
        sampler.start()
    ms, launches = timed_loop(False, timer)
    clocks = sampler.stop() if rank == 0 else None
    kern = timer.summary() if timer else {}
    ms_e2e = timed_loop(True)[0] if args.e2e else float("nan")   # --no-e2e is for profiler runs only

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    total_fl, attn_fl = flops_per_step(w)
    peak_tf, peak_gb, peak_src = measured_peaks()
    value = world * args.steps / (ms / 1e3)
    e2e_v = world * args.steps / (ms_e2e / 1e3)
    attn = kern.get("s2v_attn_fwd")
    attn_per_launch = attn_fl / (2 * w["layers"]) * (2 * P)      # 4 S^2 D per sequence, 2P sequences per launch
    roof = None
    if attn:
        ach = attn_per_launch / (attn["avg_ms"] / 1e3) / 1e12
        roof = {"kernel": "attn_fwd_kernel", "bound": "tensor", "achieved": round(ach, 1), "peak": peak_tf, "unit": "TFLOP/s",
                "frac": round(ach / peak_tf, 4),
                # dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture of this kernel at this shape
                # (profiles/r01_attn_ncu_full.txt): 706 MB + 219 MB per launch; algorithmic 705 MB (qkv) + 235 MB (out)
                "traffic": 925.2e6 if args.workload == "cfg3" else None, "algorithmic_bytes": 3 * S * D * 2 * 2 * P + S * D * 2 * 2 * P,
                "peak_source": peak_src, "avg_launch_ms": round(attn["avg_ms"], 4),
                "share_of_step": round(attn["total_ms"] / ms, 4)}
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
    kernels = {k: {"avg_ms": round(v["avg_ms"], 4), "share_of_step": round(v["total_ms"] / ms, 4),
                   **({"tflops": round(gemm_fl[k] * 2 * P / (v["avg_ms"] / 1e3) / 1e12, 1)} if k in gemm_fl else {})}
               for k, v in kern.items()}
    line = {
        "metric": "denoising_steps_per_sec", "value": round(value, 4), "unit": "steps/s (one step = one guided update of one 49f 480x720 video)",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": config_of(args, w, P_total),
        "clocks": clocks,
        "e2e": {"value": round(e2e_v, 4), "unit": "steps/s", "h2d_bytes_per_step": int(lat_h.nbytes + pe_h.nbytes + ref_h.nbytes),
                "d2h_bytes_per_step": int(out_h.nbytes), "ms_per_step": round(ms_e2e / args.steps, 3)},
        "gpu_launches": launches,
        "roofline": roof,
        "kernels": kernels,
        "model_tflops": round(total_fl * world * args.steps / (ms / 1e3) / 1e12 / world, 1),
        "build_s": round(t_build, 1),
    }
    if args.vae and w["model"] != "tiny":
        line["vae_decode"] = vae_leg(dev, lat_d, w)
    if args.cpu_baseline:
        line["cpu_baseline"] = cpu_sample(w, steps=1, warmup=0)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def vae_leg(dev, latents, w):
    """Row V next to the loop (outside the timed region): decode this rank's final latents with the 3D causal VAE
    (random-init CogVideoX decoder, the reference's default tiled schedule) and report seconds per video."""
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
    pipe = s2v_b200.CustomCogVideoXPipeline(None, None, None, vae, None)
    times = []
    for _ in range(2):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        video = pipe.decode_latents(latents[:1])
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1) / 1e3)
    assert torch.isfinite(video.float()).all()
    return {"seconds_per_video": round(min(times), 3), "schedule": "tiled 3x3 x 6 temporal batches (reference default), bf16",
            "output": list(video.shape), "conv_tflops": round(7.09e14 / min(times) / 1e12, 1)}


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
    res = cpu_sample(w, steps=max(1, args.steps), warmup=min(args.warmup, 1))
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
    ap.add_argument("--no-vae", dest="vae", action="store_false", help="skip the VAE-decode leg (reported outside the timed loop)")
    ap.add_argument("--no-e2e", dest="e2e", action="store_false", help="skip the host-buffer loop (profiler runs; not a valid bench line)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, w)
    else:
        if args.gpus > 1:
            args.cpu_baseline = False
            args.vae = False
        run_product(args, w)
