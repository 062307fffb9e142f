"""Interleaved in-process A/B of attention kernel table entries (s2v_attn_set_poly16 codes) at the cfg-3 shape.
    python tools/attn_ab.py code[:skew_ns] ...      e.g. 1:200 10:200
The order alternates between repetitions (thermal drift otherwise favours whoever runs first); medians are reported."""
import json
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from s2v_b200 import _lib, ops

B, S, H = 2, 19126, 48
torch.manual_seed(0)
qkv = torch.randn(B, S, 3 * H * 64, device="cuda").to(torch.bfloat16)
out = torch.empty(B, S, H * 64, device="cuda", dtype=torch.bfloat16)
lib = _lib.load()
cfgs = [tuple(int(x) for x in (a + ":200").split(":")[:2]) for a in sys.argv[1:]] or [(1, 200)]
reps = int(os.environ.get("REPS", "8"))


def timed(c, iters=6):
    lib.s2v_attn_set_poly16(c[0])
    lib.s2v_attn_set_skew_ns(c[1])
    for _ in range(2):
        ops.attention(qkv, out, H)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        ops.attention(qkv, out, H)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


for c in cfgs:   # warm the clocks
    timed(c)
res = {c: [] for c in cfgs}
for rep in range(reps):
    for c in (cfgs if rep % 2 == 0 else cfgs[::-1]):
        res[c].append(timed(c))
for c in cfgs:
    print(json.dumps({"code": c[0], "skew_ns": c[1], "median_ms": round(statistics.median(res[c]), 3), "ms": [round(x, 2) for x in res[c]]}))
lib.s2v_attn_set_poly16(1)
lib.s2v_attn_set_skew_ns(200)
