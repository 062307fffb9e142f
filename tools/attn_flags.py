"""In-process A/B of attention-kernel debug/tuning flags (bits of the skew word)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, json
from s2v_b200 import _lib, ops
B, S, H = 2, 19126, 48
qkv = torch.randn(B, S, 3 * H * 64, device="cuda").to(torch.bfloat16)
out = torch.empty(B, S, H * 64, device="cuda", dtype=torch.bfloat16)
lib = _lib.load()
def timed(iters=4):
    ops.attention(qkv, out, H); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): ops.attention(qkv, out, H)
    e1.record(); torch.cuda.synchronize()
    return round(e0.elapsed_time(e1) / iters, 3)
cfgs = {"base": 200, "decoupled": 200 | (1 << 26), "decoupled_noskew": (1 << 26)}
res = {k: [] for k in cfgs}
for rep in range(3):
    for k, v in cfgs.items():
        lib.s2v_attn_set_skew_ns(v); res[k].append(timed())
print(json.dumps(res))
# correctness of each flag set against the base kernel
lib.s2v_attn_set_skew_ns(200); ops.attention(qkv, out, H); ref = out.clone()
for k, v in cfgs.items():
    lib.s2v_attn_set_skew_ns(v); out.fill_(float("nan")); ops.attention(qkv, out, H); torch.cuda.synchronize()
    print(k, "max abs diff vs base", float((out.float() - ref.float()).abs().max()))
