"""Does batching more independent sequences per launch change the step's cost per video under the power governor?  The cfg-3 model forward
(42 layers; 99.9 % of a guided step) with 1, 2 and 3 prompts per GPU (CFG batch 2, 4, 6), interleaved; ms per forward per prompt, the
attention kernel's time per sequence and joules per prompt.  Not a bench line: BASELINE's configuration is one prompt per GPU."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import s2v_b200
from s2v_b200 import ops

w = bench.WORKLOADS["cfg3"]
dev = torch.device("cuda:0")
torch.cuda.set_device(0)
model, _ = bench.build_model(w, dev)
if os.environ.get("MERGE", "0") == "1":        # W + s B A folded once (two bf16 roundings away from the reference's unmerged arithmetic)
    model.merge_lora = True
    model.invalidate_engine()
n, F, S, D = bench.geometry(w)
pipe = s2v_b200.CustomCogVideoXPipeline(None, None, model, None, s2v_b200.CogVideoXDDIMScheduler.for_cogvideox(1.0))
g = torch.Generator().manual_seed(0)
rope = pipe.rotary_tables(480, 720, F, dev)
img, rr = (rope[0][n:], rope[1][n:]), (rope[0][:n], rope[1][:n])
meter = bench.EnergyMeter(0) if hasattr(bench, "EnergyMeter") else None


def inputs(P):
    lat = torch.randn(P, F, 16, 60, 90, generator=g).to(torch.bfloat16).to(dev)
    pe = (0.2 * torch.randn(2 * P, 226, 4096, generator=g)).to(torch.bfloat16).to(dev)
    ref = (0.7 * torch.randn(P, 1, 16, 60, 90, generator=g)).to(torch.bfloat16).to(dev)
    return torch.cat([lat, lat]), pe, ref          # eval=True: the reference latents are given once per prompt


def run(P, steps, data):
    x, pe, ref = data
    timer = ops.KernelTimer(["s2v_attn_fwd"])
    torch.cuda.synchronize()
    ops.set_kernel_timer(timer)
    j0 = meter.read() if meter is not None else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        model(hidden_states=x, encoder_hidden_states=pe, ref_img_states=ref, timestep=torch.full((2 * P,), 979.0, device=dev),
              image_rotary_emb=img, ref_image_rotary_emb=rr, return_dict=False, eval=True)
    e1.record()
    torch.cuda.synchronize()
    ops.set_kernel_timer(None)
    j = (meter.read() - j0) / steps / P if j0 is not None else None
    return {"ms_per_step_per_prompt": round(e0.elapsed_time(e1) / steps / P, 1),
            "attn_ms_per_sequence": round(timer.summary()["s2v_attn_fwd"]["avg_ms"] / (2 * P), 3), "joule_per_step_per_prompt": j and round(j, 1)}


Ps = [int(x) for x in os.environ.get("PROMPTS", "1,2,3").split(",")]
data = {P: inputs(P) for P in Ps}
for P in Ps:
    run(P, 1, data[P])
res = {P: [] for P in Ps}
for rep in range(int(os.environ.get("REPS", "2"))):
    for P in (Ps if rep % 2 == 0 else Ps[::-1]):
        res[P].append(run(P, int(os.environ.get("STEPS", "3")), data[P]))
for P in Ps:
    print(json.dumps({"prompts_per_gpu": P, "cfg_batch": 2 * P, "merge_lora": os.environ.get("MERGE", "0") == "1", "runs": res[P]}), flush=True)
