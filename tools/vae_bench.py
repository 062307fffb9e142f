"""Full-size VAE decode timing on one B200: 13 latent frames 60x90 (49 frames 480x720), real channel widths, random weights."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import s2v_b200
from s2v_b200 import vae as vae_mod

dev = torch.device("cuda:0")
torch.manual_seed(0)
with torch.device("meta"):
    vae = s2v_b200.AutoencoderKLCogVideoX(scaling_factor=0.7)
vae = vae.to_empty(device=dev)
g = torch.Generator(device=dev).manual_seed(0)
with torch.no_grad():
    for n, p in vae.named_parameters():
        if "norm_layer.weight" in n:
            p.copy_(1 + 0.1 * torch.randn(p.shape, device=dev, generator=g))
        elif n.endswith("bias"):
            p.copy_(0.05 * torch.randn(p.shape, device=dev, generator=g))
        else:
            fan = p[0].numel()
            p.copy_(torch.randn(p.shape, device=dev, generator=g) / fan ** 0.5)
vae = vae.to(torch.bfloat16)
z = torch.randn(1, 16, 13, 60, 90, device=dev, generator=g).to(torch.bfloat16)
for tiling in (True, False):
    if tiling:
        vae.enable_tiling(); vae.enable_slicing()
    else:
        vae.disable_tiling()
    ts = []
    flops = {}
    for it in range(3):
        vae_mod.CONV_FLOPS = flops if it == 0 else None
        torch.cuda.synchronize()
        l0 = s2v_b200._lib.launch_count
        t0 = time.time()
        if it == 2 and tiling:
            torch.cuda.nvtx.range_push("timed")   # `ncu --nvtx --nvtx-include "timed/"` captures one tiled decode
        out = vae.decode(z).sample
        if it == 2 and tiling:
            torch.cuda.nvtx.range_pop()
        torch.cuda.synchronize()
        ts.append(time.time() - t0)
        launches = s2v_b200._lib.launch_count - l0
    flop = flops["algorithmic"]      # summed over the launched convolutions (4.39e14 tiled, 3.15e14 untiled)
    print(json.dumps({"vae_decode": "tiled 3x3 (reference default)" if tiling else "untiled", "shape": list(out.shape), "finite": bool(torch.isfinite(out.float()).all()),
                      "conv_flop_algorithmic": flop, "gemm_2cta": os.environ.get("S2V_GEMM_2CTA", "default"), "seconds": [round(t, 3) for t in ts], "tflops_conv": round(flop / min(ts) / 1e12, 1), "launches": launches,
                      "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 2**30, 2)}), flush=True)
