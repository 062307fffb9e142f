"""Reference-image VAE encode and frame post-processing (row 4) timings on one B200: 480x720 image through the
5B VAE geometry (random weights), tiled as S/inference.py:206-207 sets it; uint8 conversion + device->host copy of a 49x480x720 video
against the reference's fp32 host path."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import s2v_b200
from oracle import vae_oracle as V   # parameter synthesis only (tools/ is not the product path)

dev = torch.device("cuda:0")
cfg = V.VaeConfig()
p = {k: v.to(torch.bfloat16) for k, v in V.synth_encoder_params(cfg, seed=3).items()}
m = s2v_b200.AutoencoderKLCogVideoX(scaling_factor=0.7)
m.load_state_dict(p)
m = m.to(torch.bfloat16).to(dev)
img = np.random.default_rng(0).integers(0, 256, (480, 720, 3), dtype=np.uint8)
for tiling in (True, False):
    if tiling:
        m.enable_slicing(); m.enable_tiling()
    else:
        m.disable_tiling()
    ts = []
    for _ in range(4):
        torch.cuda.synchronize()
        l0 = s2v_b200._lib.launch_count
        t0 = time.time()
        with torch.no_grad():
            lat = s2v_b200.encode_reference_image(m, img, generator=torch.Generator().manual_seed(1))
        torch.cuda.synchronize()
        ts.append(time.time() - t0)
        launches = s2v_b200._lib.launch_count - l0
    print(json.dumps({"encode_reference_image": "tiled 3x3 (reference default)" if tiling else "untiled", "out": list(lat.shape),
                      "finite": bool(torch.isfinite(lat.float()).all()), "ms": [round(1e3 * t, 2) for t in ts], "launches": launches}), flush=True)

vid = (torch.rand(1, 3, 49, 480, 720, device=dev) * 2 - 1).to(torch.bfloat16)
for name, fn in (("uint8 on device + D2H (output_type='uint8')", lambda: s2v_b200.postprocess_video(vid, "uint8")),
                 ("reference host path: fp32 frames (output_type='np') then (frame * 255).astype(uint8)",
                  lambda: (s2v_b200.postprocess_video(vid, "np") * 255).astype(np.uint8))):
    ts = []
    for _ in range(3):
        torch.cuda.synchronize()
        t0 = time.time()
        out = fn()
        ts.append(time.time() - t0)
    print(json.dumps({"postprocess": name, "shape": list(out.shape), "ms": [round(1e3 * t, 1) for t in ts]}), flush=True)
a = s2v_b200.postprocess_video(vid, "uint8")
b = (s2v_b200.postprocess_video(vid, "np") * 255).astype(np.uint8)
print(json.dumps({"uint8_paths_identical": bool(np.array_equal(a, b))}))
