"""Per-warp-slot wait profile of the attention kernel (s2v_attn_set_debug_counters) at the cfg-3 shape, isolated.
    python tools/attn_waits.py [skew_ns_word ...]      (word = skew in ns)
Prints, per setting: ms, cycles per 64-key step, and for the 8 softmax warp slots (q*4 + lane quarter) the fraction of
the tile loop spent waiting for the score tile."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from s2v_b200 import _lib, ops

B, S, H = 2, 19126, 48
torch.manual_seed(0)
qkv = torch.randn(B, S, 3 * H * 64, device="cuda").to(torch.bfloat16)
out = torch.empty(B, S, H * 64, device="cuda", dtype=torch.bfloat16)
lib = _lib.load()
dbg = torch.zeros(42, dtype=torch.int64, device="cuda")
words = [int(a, 0) for a in sys.argv[1:]] or [200, 200 | (100 << 20)]
n_cta, n_kv = B * H * ((S + 255) // 256), (S + 63) // 64
for rep in range(2):
    for w in words:
        lib.s2v_attn_set_skew_ns(w)
        lib.s2v_attn_set_debug_counters(None)
        for _ in range(3):
            ops.attention(qkv, out, H)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            ops.attention(qkv, out, H)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        dbg.zero_()
        lib.s2v_attn_set_debug_counters(dbg.data_ptr())
        ops.attention(qkv, out, H)
        torch.cuda.synchronize()
        d = [int(x) for x in dbg.tolist()]
        lib.s2v_attn_set_debug_counters(None)
        print(json.dumps({"skew_ns": w & 0xfffff, "ms": round(ms, 3), "sm_mhz": round(d[0] / max(d[1], 1) * 1e3),
                          "cycles_per_step": round(d[0] / n_cta / n_kv),
                          "wait_frac_by_slot": [round(d[2 + i] / max(d[10 + i], 1), 3) for i in range(8)],
                          "loop_cycles_per_step_by_slot": [round(d[10 + i] / n_cta / n_kv) for i in range(8)],
                          "issuer_cycles_per_step[p_ready wait, issue, kv wait] by chain": [[round(d[18 + 2 * k + i] / n_cta / n_kv) for i in range(2)] for k in range(3)]}))
lib.s2v_attn_set_skew_ns(200)
