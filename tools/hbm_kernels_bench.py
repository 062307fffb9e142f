"""Achieved HBM GB/s of the memory-bound kernels of the path at the cfg-3 shapes (CUDA events, 20 launches after warm-up; inputs
larger than L2 or an L2 flush between launches), against MEASURED_PEAKS.json's copy bandwidth.
    python tools/hbm_kernels_bench.py > gpurun_out/hbm_kernels.jsonl"""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from s2v_b200 import _lib, ops

dev = "cuda"
BF16 = torch.bfloat16
peak = 6549.8
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbps"])
except Exception:
    pass
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(iters):
        flush.zero_()                                  # evict L2 (126 MB) between launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / iters


def report(name, algorithmic_bytes, ms):
    gbps = algorithmic_bytes / ms / 1e6
    print(json.dumps({"kernel": name, "algorithmic_MB": round(algorithmic_bytes / 1e6, 1), "ms": round(ms, 4), "GB_per_s": round(gbps, 1),
                      "frac_of_measured_copy_peak": round(gbps / peak, 3), "peak_GB_per_s": peak}), flush=True)


torch.manual_seed(0)
B, S, D, H, L = 2, 19126, 3072, 48, 226
x = torch.randn(B, S, D, device=dev).to(BF16)
out = torch.empty_like(x)
w, b = (1 + 0.1 * torch.randn(D, device=dev)).to(BF16), (0.1 * torch.randn(D, device=dev)).to(BF16)
mod = torch.randn(B, 6 * D, device=dev)
report("adaln_modulate (read + write [B,S,D] bf16)", 2 * x.numel() * 2,
       timed(lambda: ops.adaln_modulate(x, out, w, b, mod, shift_off_text=3 * D, scale_off_text=4 * D, shift_off_other=0, scale_off_other=D,
                                        text_len=L, eps=1e-5)))
torch.cuda.synchronize()
print(json.dumps({"adaln_output_checksum": int(out.view(torch.int16).to(torch.int64).sum().item())}), flush=True)
qkv = torch.randn(B, S, 3 * D, device=dev).to(BF16)
nw = [(1 + 0.1 * torch.randn(64, device=dev)).to(BF16) for _ in range(2)]
nb = [(0.1 * torch.randn(64, device=dev)).to(BF16) for _ in range(2)]
ang = torch.rand(S - L, 32, device=dev) * 6.28
cos, sin = torch.cos(ang).repeat_interleave(2, 1).contiguous(), torch.sin(ang).repeat_interleave(2, 1).contiguous()
report("qk_norm_rope (read + write 2/3 of qkv; stand-alone form)", 2 * (2 * B * S * D) * 2,
       timed(lambda: ops.qk_norm_rope(qkv, nw[0], nb[0], nw[1], nb[1], cos, sin, H, L)))
fo = torch.empty(B, S - L - 1350, D, device=dev, dtype=BF16)
mod2 = torch.randn(B, 2 * D, device=dev)
report("final_norm (read video rows, write)", 2 * fo.numel() * 2,
       timed(lambda: ops.final_norm(x, fo, w, b, w, b, mod2, shift_off=0, scale_off=D, row0=L + 1350, eps=1e-5)))
vid = (torch.rand(1, 3, 49, 480, 720, device=dev) * 2 - 1).to(BF16)
report("video_to_uint8 (6 B read + 3 B written per pixel)", vid.numel() * 2 + vid.numel(), timed(lambda: ops.video_to_uint8(vid)))
# VAE elementwise kernels at the largest tiled-decode volume: 9 frames of a 240x360 tile, 128 channels
T, Hh, W, Cn, G = 9, 240, 360, 128, 32
vol = torch.randn(T + 2, Hh + 2, W + 2, Cn, device=dev).to(BF16)
vout = torch.empty_like(vol)
partial = torch.empty(1184 * Cn * 2, device=dev, dtype=torch.float32)
stats = torch.empty(G * 2, device=dev, dtype=torch.float32)
lib = _lib.load()
st = lambda: torch.cuda.current_stream().cuda_stream  # noqa: E731
report("vae groupnorm_stats (read volume)", T * Hh * W * Cn * 2,
       timed(lambda: _lib.check(lib.s2v_vae_groupnorm_stats(vol.data_ptr(), partial.data_ptr(), stats.data_ptr(), T, Hh, W, Cn, G, 1184, 1e-6, st()), "gn")))
gam, bet = torch.ones(Cn, device=dev, dtype=BF16), torch.zeros(Cn, device=dev, dtype=BF16)
report("vae groupnorm_silu (read + write volume)", 2 * T * Hh * W * Cn * 2,
       timed(lambda: _lib.check(lib.s2v_vae_groupnorm_silu(vol.data_ptr(), vout.data_ptr(), stats.data_ptr(), gam.data_ptr(), bet.data_ptr(), T, Hh, W, Cn, G, st()), "gs")))
hl, wl = Hh // 8, W // 8
yb = torch.randn(3 * hl * wl, 2 * Cn, device=dev).to(BF16)
src = (C.c_int32 * T)(*[min(t // 4, 2) for t in range(T)])
report("vae spatialnorm_silu (read + write volume)", 2 * T * Hh * W * Cn * 2,
       timed(lambda: _lib.check(lib.s2v_vae_spatialnorm_silu(vol.data_ptr(), vout.data_ptr(), stats.data_ptr(), gam.data_ptr(), bet.data_ptr(), yb.data_ptr(),
                                                             2 * Cn, src, T, Hh, W, Cn, G, hl, wl, st()), "sn")))
