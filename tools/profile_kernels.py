"""Run the dominant kernels at the cfg-3 shapes a few times (target of `ncu --set full -k regex:...`).
    python tools/profile_kernels.py attn|qkv|out|ffn_up|ffn_down|conv [iters]"""
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
