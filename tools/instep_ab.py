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


import ctypes as C

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
exp = C.CDLL(os.path.join(ROOT, "tools", "bin", "libattn_exp.so"))   # python tools/build_attn_exp.py
exp.s2v_attn_fwd_exp.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_int32, C.c_int32, C.c_int32,
                                 C.c_void_p, C.c_void_p]
exp.s2v_attn_fwd_exp.restype = C.c_int
dbg = torch.zeros(2 + 3 * 76 * 48 * 2, dtype=torch.int64, device=dev)   # sums + one (start ns, end ns, cycles) record per CTA
CUR = [0, 1, 200]
shipped_attention = ops.attention


def exp_attention(qkv, out, heads, scale=None):
    B, S_, _ = qkv.shape
    with ops._timed("s2v_attn_fwd"):
        rc = exp.s2v_attn_fwd_exp(qkv.data_ptr(), out.data_ptr(), B, S_, heads, 0.125 if scale is None else scale, CUR[0], CUR[1], CUR[2],
                                  dbg.data_ptr(), torch.cuda.current_stream().cuda_stream)
    assert rc == 0, rc
    return out


configs = {}
for a in sys.argv[1:]:   # name=variant:poly:skew   (variant bit 0 = HI warp numbering, bit 1 = K/V multicast); "shipped" = the product entry
    name, spec = a.split("=")
    configs[name] = None if spec == "shipped" else tuple(int(x) for x in spec.split(":"))
if not configs:
    configs = {"shipped": None, "r1_lo": (0, 1, 200), "hi_mc": (3, 1, 200), "bk80": (4, 1, 200), "bk80_hi_mc": (7, 1, 200)}
steps = int(os.environ.get("STEPS", "2"))
run(2)  # warm up (clocks settle on the power cap)
res = {k: [] for k in configs}
for rep in range(int(os.environ.get("REPS", "2"))):
    for k in (list(configs) if rep % 2 == 0 else list(configs)[::-1]):
        c = configs[k]
        if c is None:
            ops.attention = shipped_attention
        else:
            CUR[:] = c
            ops.attention = exp_attention
        s2v_b200.engine.ops.attention = ops.attention
        dbg.zero_()
        r = run(steps)
        cyc, ns = [int(x) for x in dbg[:2].tolist()]
        extra = ()
        if c is not None:
            # SM clock through the LAST attention launch of the run: CTAs sorted by start time, 8 equal-count bins
            rec = dbg[2:].view(-1, 3).cpu()
            rec = rec[rec[:, 2] > 0]
            rec = rec[rec[:, 0].argsort()]
            t0 = int(rec[0, 0])
            bins = []
            for ch in rec.chunk(8):
                bins.append((round(float((ch[:, 0].float().mean() - t0)) / 1e6, 2), round(float(ch[:, 2].double().sum() / (ch[:, 1] - ch[:, 0]).double().sum() * 1e3))))
            extra = (round(cyc / max(ns, 1) * 1e3), round(cyc / (steps * 42 * 7200) / 299), {"clock_mhz_through_last_launch(ms,mhz)": bins})
        res[k].append(r + extra)   # + (attention SM MHz, cycles per 64-key step, clock trajectory)
ops.attention = shipped_attention
for k, v in res.items():
    print(json.dumps({"config": k, "spec": configs[k], "attn_ms, step_ms, attn_sm_mhz, cycles_per_step": v}), flush=True)
