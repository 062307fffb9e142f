"""Run the dominant kernels at the cfg-3 shapes a few times (target of `ncu --set full -k regex:...`).
    python tools/profile_kernels.py attn|qkv|qkv_fused|out|ffn_up|ffn_down|conv|adaln|spatialnorm [iters]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import s2v_b200
from s2v_b200 import _lib, ops

which = sys.argv[1] if len(sys.argv) > 1 else "attn"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = "cuda"
torch.manual_seed(0)
B, S, H, D = 2, 19126, 48, 3072
if which == "attn":
    qkv = torch.randn(B, S, 3 * D, device=dev).to(torch.bfloat16)
    out = torch.empty(B, S, D, device=dev, dtype=torch.bfloat16)
    for _ in range(iters):
        ops.attention(qkv, out, H)
elif which == "conv":
    # the largest VAE convolution of the tiled schedule: up_block 3, 9 frames of a 240x360 tile, 128 -> 128 channels, 27 taps
    T, Hh, W, cin, cout = 9, 240, 360, 128, 128
    x = torch.randn(T + 2, Hh + 2, W + 2, cin, device=dev).to(torch.bfloat16)
    w = (torch.randn(cout, 27 * cin, device=dev) / (27 * cin) ** 0.5).to(torch.bfloat16)
    b = torch.zeros(cout, device=dev, dtype=torch.bfloat16)
    o = torch.empty(T + 2, Hh + 2, W + 2, cout, device=dev, dtype=torch.bfloat16)
    a = _lib.ConvArgs()
    a.x, a.ldx, a.w, a.ldw, a.bias, a.res, a.ldres, a.out, a.ldo = x.data_ptr(), cin, w.data_ptr(), 27 * cin, b.data_ptr(), None, 0, o.data_ptr(), cout
    a.T, a.t_pad, a.Hp, a.Wp, a.cin, a.cout, a.taps = T, 2, Hh + 2, W + 2, cin, cout, 27
    for _ in range(iters):
        _lib.check(_lib.load().s2v_conv_gemm(C.byref(a), torch.cuda.current_stream().cuda_stream), "conv")
elif which == "adaln":
    x = torch.randn(B, S, D, device=dev).to(torch.bfloat16)
    o = torch.empty_like(x)
    w, b = torch.ones(D, device=dev, dtype=torch.bfloat16), torch.zeros(D, device=dev, dtype=torch.bfloat16)
    mod = torch.randn(B, 6 * D, device=dev)
    for _ in range(iters):
        ops.adaln_modulate(x, o, w, b, mod, shift_off_text=3 * D, scale_off_text=4 * D, shift_off_other=0, scale_off_other=D, text_len=226, eps=1e-5)
elif which == "spatialnorm":
    # decoder SpatialNorm + SiLU at the largest tiled-decode volume: 9 frames of a 240x360 tile, 128 channels
    T, Hh, W, Cn, G = 9, 240, 360, 128, 32
    vol = torch.randn(T + 2, Hh + 2, W + 2, Cn, device=dev).to(torch.bfloat16)
    vout = torch.empty_like(vol)
    partial = torch.empty(1184 * Cn * 2, device=dev, dtype=torch.float32)
    stats = torch.empty(G * 2, device=dev, dtype=torch.float32)
    gam, bet = torch.ones(Cn, device=dev, dtype=torch.bfloat16), torch.zeros(Cn, device=dev, dtype=torch.bfloat16)
    hl, wl = Hh // 8, W // 8
    yb = torch.randn(3 * hl * wl, 2 * Cn, device=dev).to(torch.bfloat16)
    src = (C.c_int32 * T)(*[min(t // 4, 2) for t in range(T)])
    lib, st = _lib.load(), torch.cuda.current_stream().cuda_stream
    for _ in range(iters):
        _lib.check(lib.s2v_vae_groupnorm_stats(vol.data_ptr(), partial.data_ptr(), stats.data_ptr(), T, Hh, W, Cn, G, 1184, 1e-6, st), "gn")
        _lib.check(lib.s2v_vae_spatialnorm_silu(vol.data_ptr(), vout.data_ptr(), stats.data_ptr(), gam.data_ptr(), bet.data_ptr(), yb.data_ptr(),
                                                2 * Cn, src, T, Hh, W, Cn, G, hl, wl, st), "sn")
elif which == "qkv_fused":
    M, H, L = B * S, 48, 226
    x = torch.randn(M, D, device=dev).to(torch.bfloat16)
    w = (0.02 * torch.randn(3 * D, D, device=dev)).to(torch.bfloat16)
    b = (0.02 * torch.randn(3 * D, device=dev)).to(torch.bfloat16)
    t = torch.randn(M, 3 * 128, device=dev).to(torch.bfloat16)
    lb = (0.02 * torch.randn(3 * D, 128, device=dev)).to(torch.bfloat16)
    nw = [torch.ones(64, device=dev, dtype=torch.bfloat16) for _ in range(2)]
    nb = [torch.zeros(64, device=dev, dtype=torch.bfloat16) for _ in range(2)]
    ang = torch.rand(S - L, 32, device=dev) * 6.28
    cos, sin = torch.cos(ang).repeat_interleave(2, 1).contiguous(), torch.sin(ang).repeat_interleave(2, 1).contiguous()
    o = torch.empty(M, 3 * D, device=dev, dtype=torch.bfloat16)
    qk = ops.qk_norm_args(nw[0], nb[0], nw[1], nb[1], cos, sin, S, H, L)
    for _ in range(iters):
        ops.linear(x, w, b, o, lora_t=t, lora_b=lb, lora_group_n=D, entry="s2v_qkv_lora", qk=qk)
else:
    M = B * S
    shapes = {"qkv": (3 * D, D), "out": (D, D), "ffn_up": (4 * D, D), "ffn_down": (D, 4 * D)}
    N, K = shapes[which]
    x = torch.randn(M, K, device=dev).to(torch.bfloat16)
    w = (0.02 * torch.randn(N, K, device=dev)).to(torch.bfloat16)
    b = (0.02 * torch.randn(N, device=dev)).to(torch.bfloat16)
    o = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    for _ in range(iters):
        ops.linear(x, w, b, o)
torch.cuda.synchronize()
print("done", which)
