"""In-step A/B of attention tuning knobs under the real power regime: the full cfg-3 guided step (42 layers, GEMMs +
attention + elementwise) is run for a few steps per setting, interleaved, and the attention kernel's average CUDA-event
time and the step time are reported.  Isolated sweeps mislead here: inside the step the board sits on its 1 kW cap."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import s2v_b200
from s2v_b200 import _lib, ops

w = bench.WORKLOADS["cfg3"]
dev = torch.device("cuda:0")
torch.cuda.set_device(0)
model, _ = bench.build_model(w, dev)
n, F, S, D = bench.geometry(w)
sched = s2v_b200.CogVideoXDDIMScheduler.for_cogvideox(1.0)
sched.set_timesteps(50)
pipe = s2v_b200.CustomCogVideoXPipeline(None, None, model, None, sched)
g = torch.Generator().manual_seed(0)
lat = torch.randn(1, F, 16, 60, 90, generator=g).to(torch.bfloat16).to(dev)
pe = (0.2 * torch.randn(2, 226, 4096, generator=g)).to(torch.bfloat16).to(dev)
ref = (0.7 * torch.randn(1, 1, 16, 60, 90, generator=g)).to(torch.bfloat16).to(dev)
rope = pipe.rotary_tables(480, 720, F, dev)
img, rr = (rope[0][n:], rope[1][n:]), (rope[0][:n], rope[1][:n])
model_in = torch.cat([lat, lat])
lib = _lib.load()


def run(steps):
    timer = ops.KernelTimer(["s2v_attn_fwd"])
    torch.cuda.synchronize()
    ops.set_kernel_timer(timer)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        model(hidden_states=model_in, encoder_hidden_states=pe, ref_img_states=ref, timestep=torch.full((2,), 979.0, device=dev),
              image_rotary_emb=img, ref_image_rotary_emb=rr, return_dict=False, eval=True)
    e1.record()
    torch.cuda.synchronize()
    ops.set_kernel_timer(None)
    return round(timer.summary()["s2v_attn_fwd"]["avg_ms"], 3), round(e0.elapsed_time(e1) / steps, 1)


configs = {}
for a in sys.argv[1:]:
    name, variant, poly, skew = (a.split(":") + ["default", "1", "200"])[:4] if a.count(":") == 0 else a.split(":")
    configs[name] = (variant, int(poly), int(skew))
if not configs:
    configs = {"default_p1_s200": ("default", 1, 200), "default_p0_s200": ("default", 0, 200), "default_p2_s200": ("default", 2, 200),
               "default_p1_s0": ("default", 1, 0), "v4_p1_s0": ("v4", 1, 0)}
dbg = torch.zeros(42, dtype=torch.int64, device=dev)
lib.s2v_attn_set_debug_counters(dbg.data_ptr())
run(2)  # warm up (clocks settle on the power cap)
res = {k: [] for k in configs}
for rep in range(2):
    for k, (variant, poly, skew) in configs.items():
        ops.ATTN_VARIANT = variant
        ops.ATTN_V4_POLY16, ops.ATTN_V4_SKEW_NS = poly, skew
        lib.s2v_attn_set_poly16(poly)
        lib.s2v_attn_set_skew_ns(skew)
        dbg.zero_()
        r = run(2)
        c, ns = [int(x) for x in dbg.tolist()[:2]]
        res[k].append(r + (round(c / max(ns, 1) * 1e3), round(c / (2 * 42 * 7200) / 299)))   # + (attention SM MHz, cycles per 64-key step)
for k, v in res.items():
    print(json.dumps({"config": k, "attn_ms, step_ms, attn_sm_mhz, cycles_per_step": v}))

# the same kernel in isolation (sustained loop), for the clock / cycle comparison
qkv = torch.randn(2, S, 3 * D, device=dev).to(torch.bfloat16)
out = torch.empty(2, S, D, device=dev, dtype=torch.bfloat16)
lib.s2v_attn_set_poly16(1); lib.s2v_attn_set_skew_ns(200); ops.ATTN_VARIANT = "default"
for _ in range(20): ops.attention(qkv, out, w["heads"])
torch.cuda.synchronize(); dbg.zero_()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): ops.attention(qkv, out, w["heads"])
e1.record(); torch.cuda.synchronize()
c, ns = [int(x) for x in dbg.tolist()[:2]]
print(json.dumps({"isolated_sustained_attn_ms": round(e0.elapsed_time(e1) / 20, 3), "attn_sm_mhz": round(c / ns * 1e3), "cycles_per_step": round(c / (20 * 7200) / 299)}))
