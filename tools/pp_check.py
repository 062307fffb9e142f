import sys, os
sys.path.insert(0, "/root/repo")
import torch
from s2v_b200 import _lib, ops
lib = _lib.load()
torch.manual_seed(0)
B, S, H = 1, 1500, 2
qkv = torch.randn(B, S, 3*H*64, device="cuda").to(torch.bfloat16)
o0 = torch.empty(B, S, H*64, device="cuda", dtype=torch.bfloat16); o1 = torch.empty_like(o0)
lib.s2v_attn_set_skew_ns(200); ops.attention(qkv, o0, H)
lib.s2v_attn_set_skew_ns(1 << 28); ops.attention(qkv, o1, H)
torch.cuda.synchronize()
print("pp max diff", (o0.float()-o1.float()).abs().max().item(), "equal", torch.equal(o0, o1))
