"""Per-phase timeline of the attention kernel's softmax warps and issuing thread (ATT_TRACE build of the experiment library):
CTA (1,0,0) of one cfg-3 launch, tiles 64..79.  Prints, per softmax warp, the mean cycles between consecutive stamps
  wait_S (s_full wait) | ld (tcgen05.ld + wait) | compute (exponentials .. pack) | st (tcgen05.st + wait::st) | publish
and the tile period; for the issuing thread the time from a warp group's P being published to its PV issue and the issue duration."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

B, S, H = 2, 19126, 48
torch.manual_seed(0)
qkv = torch.randn(B, S, 3 * H * 64, device="cuda").to(torch.bfloat16)
out = torch.empty(B, S, H * 64, device="cuda", dtype=torch.bfloat16)
exp = C.CDLL(os.path.join(ROOT, "tools", "bin", "libattn_exp.so"))
vp, i32 = C.c_void_p, C.c_int32
exp.s2v_attn_fwd_exp.argtypes = [vp, vp, i32, i32, i32, C.c_float, i32, i32, i32, vp, vp]
OFF, NT, NS = 2 + 3 * 76 * 48 * 2, 16, 6
dbg = torch.zeros(OFF + 8 * NT * NS + 2 * NT * 2 + 16, dtype=torch.int64, device="cuda")
variant = int(sys.argv[1]) if len(sys.argv) > 1 else 16
for _ in range(3):
    dbg.zero_()
    rc = exp.s2v_attn_fwd_exp(qkv.data_ptr(), out.data_ptr(), B, S, H, 0.125, variant, 1, 200, dbg.data_ptr(), torch.cuda.current_stream().cuda_stream)
    assert rc == 0, rc
    torch.cuda.synchronize()
d = dbg.cpu()
st = d[OFF:OFF + 8 * NT * NS].view(8, NT, NS).to(torch.int64) & 0xFFFFFFFF
iss = d[OFF + 8 * NT * NS:OFF + 8 * NT * NS + 2 * NT * 2].view(2, NT, 2) & 0xFFFFFFFF
names = ["wait_S", "ld", "compute", "st", "publish"]
for w in range(8):
    x = st[w]
    seg = ((x[:, 1:] - x[:, :-1]) & 0xFFFFFFFF).float().mean(0).tolist()
    period = float(((x[1:, 0] - x[:-1, 0]) & 0xFFFFFFFF).float().mean())
    print(json.dumps({"warp": w, "q": w // 4, "smsp": w % 4, **{n: round(v) for n, v in zip(names, seg)}, "tile_period": round(period)}))
for q in range(2):
    pub = st[4 * q:4 * q + 4, :, 5].max(0).values          # last of the 4 warps' publish stamps
    lag = ((iss[q, :, 0] - pub) & 0xFFFFFFFF).float()
    lag = torch.where(lag > 2 ** 31, lag - 2 ** 32, lag)
    dur = ((iss[q, :, 1] - iss[q, :, 0]) & 0xFFFFFFFF).float()
    print(json.dumps({"issuer_chain": q, "publish_to_pv_issue": [round(v) for v in lag.tolist()], "pv_plus_s_issue_cycles": round(float(dur.mean()))}))
t0 = int(st[:, 0, 0].min())
print("timeline (cycles from the first stamp; warp: start, S ready, loaded, computed, stored, published) for tiles 64..67")
for t in range(4):
    for w in range(8):
        print(t + 64, w, [int((int(v) - t0) & 0xFFFFFFFF) for v in st[w, t].tolist()])
