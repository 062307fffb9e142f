import sys, os
sys.path.insert(0, "/root/repo")
os.environ["S2V_ATTN_VARIANT"] = "v4"
import torch
from s2v_b200 import ops
B, S, H, D = 2, 19126, 48, 3072
torch.manual_seed(0)
qkv = torch.randn(B, S, 3 * D, device="cuda").to(torch.bfloat16)
out = torch.empty(B, S, D, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    ops.attention(qkv, out, H)
torch.cuda.synchronize()
