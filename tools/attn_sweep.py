"""Sweep the attention kernel's polynomial-exp2 fraction at the cfg-3 shape, interleaved with torch SDPA as a clock reference."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

from s2v_b200 import _lib, ops

B, S, H = 2, 19126, 48
torch.manual_seed(0)
qkv = torch.randn(B, S, 3 * H * 64, device="cuda").to(torch.bfloat16)
out = torch.empty(B, S, H * 64, device="cuda", dtype=torch.bfloat16)
q, k, v = [t.view(B, S, H, 64).transpose(1, 2) for t in qkv.chunk(3, dim=-1)]
fl = 4.0 * B * H * S * S * 64


def timed(fn, iters=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


lib = _lib.load()
# usage: attn_sweep.py poly16[:skew_ns] ...
vals = [tuple(int(x) for x in (a + ":0").split(":")[:2]) for a in sys.argv[1:]] or [(p, 0) for p in (0, 1, 2, 3, 4)]
res = {p: [] for p in vals}
ref = []
for rep in range(3):
    ref.append(timed(lambda: F.scaled_dot_product_attention(q, k, v)))
    for p in vals:
        lib.s2v_attn_set_poly16(p[0])
        lib.s2v_attn_set_skew_ns(p[1])
        res[p].append(timed(lambda: ops.attention(qkv, out, H)))
def v4(poly=1, skew=200):
    _lib.check(lib.s2v_attn_fwd_v4(qkv.data_ptr(), out.data_ptr(), B, S, H, 0.125, poly, skew, torch.cuda.current_stream().cuda_stream), "v4")


def v1():
    _lib.check(lib.s2v_attn_fwd_v1(qkv.data_ptr(), out.data_ptr(), B, S, H, 0.125, torch.cuda.current_stream().cuda_stream), "v1")


if hasattr(lib, "s2v_attn_fwd_v1") and "v1" in os.environ.get("S2V_SWEEP", "v1"):
    lib.s2v_attn_set_poly16(1); lib.s2v_attn_set_skew_ns(200)
    ops.attention(qkv, out, H); o_main = out.clone(); v1(); torch.cuda.synchronize()
    print(json.dumps({"v1_max_abs_diff_vs_main": float((out.float() - o_main.float()).abs().max())}))
    a, b = [], []
    for rep in range(4):
        a.append(timed(lambda: ops.attention(qkv, out, H)))
        b.append(timed(v1))
    print(json.dumps({"main_ms": [round(x, 3) for x in a], "v1_elect_ms": [round(x, 3) for x in b]}))
if hasattr(lib, "s2v_attn_fwd_v4") and "v4" in os.environ.get("S2V_SWEEP", ""):
    lib.s2v_attn_set_poly16(1); lib.s2v_attn_set_skew_ns(200)
    ops.attention(qkv, out, H); o_main = out.clone(); v4(); torch.cuda.synchronize()
    print(json.dumps({"v4_max_abs_diff_vs_main": float((out.float() - o_main.float()).abs().max())}))
    cfgs = [(0, 0), (1, 0), (1, 200), (2, 200), (1, 400)]
    a, b = [], {c: [] for c in cfgs}
    for rep in range(3):
        a.append(timed(lambda: ops.attention(qkv, out, H)))
        for c in cfgs:
            b[c].append(timed(lambda: v4(*c)))
    print(json.dumps({"main_ms": [round(x, 3) for x in a]}))
    for c in cfgs:
        print(json.dumps({"v4 poly16,skew": c, "ms": [round(x, 3) for x in b[c]]}))
print(json.dumps({"torch_sdpa_ms": [round(x, 3) for x in ref], "tflops": round(fl / min(ref) / 1e9, 1)}))
for p in vals:
    print(json.dumps({"poly16": p[0], "skew_ns": p[1], "ms": [round(x, 3) for x in res[p]], "best_tflops": round(fl / min(res[p]) / 1e9, 1)}))
