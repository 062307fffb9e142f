"""SM clock trajectory through a full cfg-3 denoising step (tools/clock_probe.cu): a one-thread probe kernel samples
(%globaltimer, clock64) every 20 us on its own stream while the step runs; NVTX-free phase labels come from CUDA events recorded
around every attention launch and every big GEMM.  Prints, per kernel class, the mean clock and the clock at the start / middle /
end of the launches, and the clock trajectory through one layer — the data behind "why the step time does not follow the kernels'
isolated times" in profiles/r02_summary.md.
    S2V_GEMM_2CTA=0|1 python tools/clock_trace.py"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import bench
import s2v_b200
from s2v_b200 import ops

w = bench.WORKLOADS["cfg3"]
dev = torch.device("cuda:0")
torch.cuda.set_device(0)
model, _ = bench.build_model(w, dev)
n, F, S, D = bench.geometry(w)
sched = s2v_b200.CogVideoXDDIMScheduler.for_cogvideox(1.0)
sched.set_timesteps(50)
ALL_STEPS = list(sched._timesteps_host)   # the pipeline call below installs its own (shorter) list on the scheduler
pipe = s2v_b200.CustomCogVideoXPipeline(None, None, model, None, sched)
inp = bench.Inputs(w, 1, dev, 0, 4096)
exp = C.CDLL(os.path.join(ROOT, "tools", "bin", "libattn_exp.so"))
exp.s2v_clock_probe.argtypes = [C.c_void_p, C.c_int32, C.c_uint32, C.c_void_p, C.c_void_p]


def steps(k, first=0):
    ts = [ALL_STEPS[(first + i) % 50] for i in range(k)]
    return pipe(prompt=None, prompt_embeds=inp.pos_d, negative_prompt_embeds=inp.neg_d, ref_img_states=inp.ref_d, latents=inp.lat_d, height=480,
                width=720, num_frames=49, num_inference_steps=50, timesteps=ts, guidance_scale=6.0, output_type="latent", return_dict=False)[0]


steps(3)            # settle on the power regime
torch.cuda.synchronize()
N = 120000          # 2.4 s at 20 us
buf = torch.zeros(2 * N, dtype=torch.int64, device=dev)
stop = torch.zeros(1, dtype=torch.int32, device=dev)
probe_stream = torch.cuda.Stream(priority=-1)
names = ["s2v_attn_fwd", "s2v_qkv_lora", "s2v_outproj_lora_gate_residual", "s2v_ffn_up_gelu_lora", "s2v_ffn_down_lora_gate_residual"]
timer = ops.KernelTimer(names)
with torch.cuda.stream(probe_stream):
    exp.s2v_clock_probe(buf.data_ptr(), N, 20000, stop.data_ptr(), probe_stream.cuda_stream)
ref = torch.cuda.Event(enable_timing=True)
ref.record()
ops.set_kernel_timer(timer)
steps(2, 3)
ops.set_kernel_timer(None)
torch.cuda.current_stream().synchronize()
stop.fill_(1)
torch.cuda.synchronize()
s = buf.view(-1, 2).cpu()
s = s[s[:, 0] > 0]
t_ns, cyc = s[:, 0].double(), s[:, 1].double()
mid = ((t_ns[1:] + t_ns[:-1]) / 2 - t_ns[0]) / 1e6                 # ms since the first sample
mhz = (cyc[1:] - cyc[:-1]) / (t_ns[1:] - t_ns[:-1]) * 1e3
# kernel intervals relative to `ref` (recorded right after the probe started; the probe's first sample is ~the same instant)
out = {"samples": int(len(mhz)), "span_ms": round(float(mid[-1]), 1), "clock_mhz_overall_mean": round(float(mhz.mean()))}
per = {}
for name, evs in timer.events.items():
    rows = []
    for a, b in evs:
        t0, t1 = ref.elapsed_time(a), ref.elapsed_time(b)
        sel = (mid >= t0) & (mid <= t1)
        if sel.sum() >= 3:
            m = mhz[sel]
            k = len(m)
            rows.append((float(m.mean()), float(m[: max(k // 5, 1)].mean()), float(m[k // 2 - max(k // 10, 1): k // 2 + max(k // 10, 1)].mean()),
                         float(m[-max(k // 5, 1):].mean()), t1 - t0))
    if rows:
        t = torch.tensor(rows)
        per[name] = {"launches": len(rows), "ms": round(float(t[:, 4].mean()), 3), "clock_mean": round(float(t[:, 0].mean())),
                     "clock_first_fifth": round(float(t[:, 1].mean())), "clock_middle": round(float(t[:, 2].mean())),
                     "clock_last_fifth": round(float(t[:, 3].mean()))}
out["per_kernel"] = per
# one layer in the middle of the second step, 0.5 ms bins
a0 = timer.events["s2v_attn_fwd"][60][0]
a1 = timer.events["s2v_attn_fwd"][61][0]
t0, t1 = ref.elapsed_time(a0), ref.elapsed_time(a1)
layer = []
x = t0
while x < t1:
    sel = (mid >= x) & (mid < x + 0.5)
    if sel.any():
        layer.append((round(x - t0, 1), round(float(mhz[sel].mean()))))
    x += 0.5
out["one_layer_clock(ms_from_attention_start,mhz)"] = layer
out["gemm_2cta"] = os.environ.get("S2V_GEMM_2CTA", "default")
print(json.dumps(out), flush=True)
