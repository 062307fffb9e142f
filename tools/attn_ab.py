"""Interleaved in-process A/B of attention kernel variants at the cfg-3 shape (B=2, S=19126, H=48), isolated (sustained loop),
next to torch SDPA (cuDNN / flash backend) on the same tensors as the library reference point.

    python tools/build_attn_exp.py            # here (cross-compile)
    python tools/attn_ab.py [name=variant:poly:skew ...]      # on the GPU box; variant bit0 = HI warp numbering, bit1 = K/V multicast, bit2 = 80-key tiles
The order alternates between repetitions (thermal drift otherwise favours whoever runs first); medians are reported, and every
variant's output is checked against the shipped kernel's (max abs difference) before it is timed."""
import ctypes as C
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.nn.functional as F

from s2v_b200 import ops

B, S, H = (int(x) for x in os.environ.get("SHAPE", "2,19126,48").split(","))
torch.manual_seed(0)
qkv = torch.randn(B, S, 3 * H * 64, device="cuda").to(torch.bfloat16)
out = torch.empty(B, S, H * 64, device="cuda", dtype=torch.bfloat16)
exp = C.CDLL(os.path.join(ROOT, "tools", "bin", "libattn_exp.so"))
vp, i32 = C.c_void_p, C.c_int32
exp.s2v_attn_fwd_exp.argtypes = [vp, vp, i32, i32, i32, C.c_float, i32, i32, i32, vp, vp]
exp.s2v_attn_fwd_exp.restype = C.c_int
exp.s2v_last_error.restype = C.c_char_p
dbg = torch.zeros(2 + 3 * 76 * 48 * 2, dtype=torch.int64, device="cuda")
reps = int(os.environ.get("REPS", "6"))
iters = int(os.environ.get("ITERS", "6"))

cfgs = {}
for a in sys.argv[1:]:
    name, spec = a.split("=")
    cfgs[name] = tuple(int(x) for x in spec.split(":"))
if not cfgs:
    cfgs = {"r1_lo": (0, 1, 200), "hi_mc": (3, 1, 200), "bk80": (4, 1, 200), "bk80_hi": (5, 1, 200), "bk80_mc": (6, 1, 200),
            "bk80_hi_mc": (7, 1, 200), "bk80_hi_mc_p0": (7, 0, 200), "bk80_hi_mc_p2": (7, 2, 200), "bk80_hi_mc_s0": (7, 1, 0),
            "bk80_hi_mc_s400": (7, 1, 400)}


USE_DBG = os.environ.get("DBG", "1") == "1"      # DBG=0: no per-CTA counters (thread 0's clock reads delay softmax warp 0 at the start)


def run_exp(c, o):
    rc = exp.s2v_attn_fwd_exp(qkv.data_ptr(), o.data_ptr(), B, S, H, 0.125, c[0], c[1], c[2], dbg.data_ptr() if USE_DBG else None,
                              torch.cuda.current_stream().cuda_stream)
    if rc:
        raise RuntimeError(f"s2v_attn_fwd_exp rc={rc}: {exp.s2v_last_error().decode()}")


q4 = qkv.view(B, S, 3, H, 64)
qs, ks, vs = (q4[:, :, i].transpose(1, 2) for i in range(3))   # [B,H,S,64] strided views, like the reference's transposes


def run_sdpa(_c, _o):
    return F.scaled_dot_product_attention(qs, ks, vs)


def run_shipped(_c, o):
    ops.attention(qkv, o, H)


def timed(fn, c):
    for _ in range(2):
        fn(c, out)
    torch.cuda.synchronize()
    dbg.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn(c, out)
    e1.record()
    torch.cuda.synchronize()
    cyc, ns = (int(x) for x in dbg[:2].tolist())
    return e0.elapsed_time(e1) / iters, (cyc / ns * 1e3 if ns else 0.0)


# correctness of every variant against the shipped entry point (and of that against SDPA)
ref = torch.empty_like(out)
ops.attention(qkv, ref, H)
sd = run_sdpa(None, None).transpose(1, 2).reshape(B, S, H * 64)
torch.cuda.synchronize()
print(json.dumps({"shipped_vs_sdpa_max_abs": float((ref.float() - sd.float()).abs().max())}), flush=True)
for name, c in cfgs.items():
    o = torch.full_like(out, float("nan"))
    run_exp(c, o)
    torch.cuda.synchronize()
    print(json.dumps({"variant": name, "spec": c, "max_abs_vs_shipped": float((o.float() - ref.float()).abs().max())}), flush=True)

arms = {"sdpa": (run_sdpa, None), "shipped": (run_shipped, None)}
_old_path = os.path.join(ROOT, "tools", "bin", "libs2v_old.so")      # optional: a previous build of the product library as its own arm
if os.path.exists(_old_path) and os.environ.get("OLD", "0") == "1":
    _old = C.CDLL(_old_path)
    _old.s2v_attn_fwd.argtypes = [vp, vp, i32, i32, i32, C.c_float, vp]

    def run_old(_c, o):
        assert _old.s2v_attn_fwd(qkv.data_ptr(), o.data_ptr(), B, S, H, 0.125, torch.cuda.current_stream().cuda_stream) == 0

    arms["old_build"] = (run_old, None)
arms.update({k: (run_exp, c) for k, c in cfgs.items()})
for k, (fn, c) in arms.items():   # warm the clocks
    timed(fn, c)
res = {k: [] for k in arms}
order = list(arms)
for rep in range(reps):
    for k in (order if rep % 2 == 0 else order[::-1]):
        res[k].append(timed(*arms[k]))
fl = 4.0 * S * S * 64 * H * B
for k in arms:
    ms = statistics.median(x[0] for x in res[k])
    mhz = statistics.median(x[1] for x in res[k])
    print(json.dumps({"arm": k, "median_ms": round(ms, 3), "tflops": round(fl / ms / 1e9, 1), "sm_mhz": round(mhz),
                      "ms": [round(x[0], 2) for x in res[k]]}), flush=True)
