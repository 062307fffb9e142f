import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from s2v_b200 import ops
B, S, H, scale = [float(x) if "." in x else int(x) for x in sys.argv[1:5]]
torch.manual_seed(0)
qkv = (torch.randn(B, S, 3 * H * 64, device="cuda") * scale).to(torch.bfloat16)
out = torch.full((B, S, H * 64), float("nan"), device="cuda", dtype=torch.bfloat16)
ops.attention(qkv, out, H)
torch.cuda.synchronize()
q, k, v = [t.view(B, S, H, 64).transpose(1, 2).float() for t in qkv.chunk(3, dim=-1)]
ref = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B, S, H * 64)
print("case", sys.argv[1:5], "max err", float((out.float() - ref).abs().max()), "ref max", float(ref.abs().max()))
