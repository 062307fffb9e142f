"""Run the dominant kernels at the cfg-3 shapes a few times (target of `ncu --set full -k regex:...`)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import s2v_b200
from s2v_b200 import ops

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
