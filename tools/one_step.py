"""Two cfg-3 model forwards (target of an ncu launch list: `ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip <launches of
the first forward> --csv --log-file out.csv python tools/one_step.py`; MERGE=1 folds the LoRA factors into the weights)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import s2v_b200
from s2v_b200 import _lib

w = bench.WORKLOADS["cfg3"]
dev = torch.device("cuda:0")
model, _ = bench.build_model(w, dev)
if os.environ.get("MERGE", "0") == "1":
    model.merge_lora = True
    model.invalidate_engine()
n, F, S, D = bench.geometry(w)
pipe = s2v_b200.CustomCogVideoXPipeline(None, None, model, None, s2v_b200.CogVideoXDDIMScheduler.for_cogvideox(1.0))
g = torch.Generator().manual_seed(0)
rope = pipe.rotary_tables(480, 720, F, dev)
img, rr = (rope[0][n:], rope[1][n:]), (rope[0][:n], rope[1][:n])
lat = torch.randn(1, F, 16, 60, 90, generator=g).to(torch.bfloat16).to(dev)
pe = (0.2 * torch.randn(2, 226, 4096, generator=g)).to(torch.bfloat16).to(dev)
ref = (0.7 * torch.randn(1, 1, 16, 60, 90, generator=g)).to(torch.bfloat16).to(dev)
x = torch.cat([lat, lat])
for i in range(int(os.environ.get("FORWARDS", "2"))):
    c0 = _lib.launch_count
    model(hidden_states=x, encoder_hidden_states=pe, ref_img_states=ref, timestep=torch.full((2,), 979.0, device=dev),
          image_rotary_emb=img, ref_image_rotary_emb=rr, return_dict=False, eval=True)
    torch.cuda.synchronize()
    print("forward", i, "launches through the C ABI:", _lib.launch_count - c0, flush=True)
