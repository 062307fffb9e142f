import sys, os, json, statistics
sys.path.insert(0, os.getcwd())
import torch
from s2v_b200 import ops
dev="cuda"; B,S,H,D,L=2,19126,48,3072,226
M=B*S
torch.manual_seed(0)
x = torch.randn(M, D, device=dev).to(torch.bfloat16)
w = (0.02 * torch.randn(3 * D, D, device=dev)).to(torch.bfloat16)
b = (0.02 * torch.randn(3 * D, device=dev)).to(torch.bfloat16)
t = torch.randn(M, 3 * 128, device=dev).to(torch.bfloat16)
lb = (0.02 * torch.randn(3 * D, 128, device=dev)).to(torch.bfloat16)
nw = [(1+0.1*torch.randn(64, device=dev)).to(torch.bfloat16) for _ in range(2)]
nb = [(0.1*torch.randn(64, device=dev)).to(torch.bfloat16) for _ in range(2)]
ang = torch.rand(S - L, 32, device=dev) * 6.28
cos, sin = torch.cos(ang).repeat_interleave(2, 1).contiguous(), torch.sin(ang).repeat_interleave(2, 1).contiguous()
o = torch.empty(M, 3 * D, device=dev, dtype=torch.bfloat16)
o2 = torch.empty_like(o)
qk = ops.qk_norm_args(nw[0], nb[0], nw[1], nb[1], cos, sin, S, H, L)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def fused(): ops.linear(x, w, b, o, lora_t=t, lora_b=lb, lora_group_n=D, entry="s2v_qkv_lora", qk=qk)
def plain(): ops.linear(x, w, b, o2, lora_t=t, lora_b=lb, lora_group_n=D, entry="s2v_qkv_lora")
def timed(fn, n=15):
    for _ in range(3): fn()
    torch.cuda.synchronize(); ms=[]
    for _ in range(n):
        flush.zero_()
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ms.append(e0.elapsed_time(e1))
    return statistics.median(ms)
tf, tp = timed(fused), timed(plain)
ops.qk_norm_rope(o2.view(B,S,3*D), nw[0], nb[0], nw[1], nb[1], cos, sin, H, L)
torch.cuda.synchronize()
print(json.dumps({"qkv_fused_ms": round(tf,4), "qkv_plain_ms": round(tp,4), "fused_equals_plain_plus_standalone": bool(torch.equal(o, o2))}))
