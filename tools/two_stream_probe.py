"""Feasibility probe for evening out the step's power density in time (profiles/r02_summary.md §5): the kernel mix of one cfg-3
layer (QKV GEMM, attention, out-proj, FFN-up, FFN-down; LoRA and the small kernels left out) for TWO sequences, either
  serial   B = 2 batched in every launch, one stream (what the engine does), or
  split    one sequence per stream pair, GEMMs on a high-priority stream confined to S2V_GEMM_SMS persistent CTAs, attention on a
           low-priority stream filling the remaining SMs, the second sequence half a layer behind the first.
Prints ms per layer (both sequences) and joules per layer for both schedules.
    S2V_GEMM_SMS=52 python tools/two_stream_probe.py split ;  python tools/two_stream_probe.py serial"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import bench
from s2v_b200 import ops

mode = sys.argv[1] if len(sys.argv) > 1 else "serial"
LAYERS = int(os.environ.get("LAYERS", "42"))
REPS = int(os.environ.get("REPS", "3"))
dev = torch.device("cuda:0")
S, D, H = 19126, 3072, 48
bf = torch.bfloat16
torch.manual_seed(0)


def make(B):
    M = B * S
    t = dict(x=torch.randn(M, D, device=dev).to(bf), qkv=torch.empty(B, S, 3 * D, device=dev, dtype=bf), att=torch.empty(B, S, D, device=dev, dtype=bf),
             h=torch.randn(M, D, device=dev).to(bf), ffh=torch.empty(M, 4 * D, device=dev, dtype=bf), mod=torch.randn(B, 6 * D, device=dev))
    return t


W = dict(qkv=(0.02 * torch.randn(3 * D, D, device=dev)).to(bf), out=(0.02 * torch.randn(D, D, device=dev)).to(bf),
         up=(0.02 * torch.randn(4 * D, D, device=dev)).to(bf), down=(0.02 * torch.randn(D, 4 * D, device=dev)).to(bf))
bias = {k: torch.zeros(v.shape[0], device=dev, dtype=bf) for k, v in W.items()}
gate = dict(epilogue=ops.EPI_GATE_RESIDUAL, gate_off_text=5 * D, gate_off_other=2 * D, rows_per_batch=S, text_len=226)


def pre_attn(t):
    ops.linear(t["x"], W["qkv"], bias["qkv"], t["qkv"].view(-1, 3 * D))


def attn(t):
    ops.attention(t["qkv"], t["att"], H)


def post_attn(t):
    ops.linear(t["att"].view(-1, D), W["out"], bias["out"], t["h"], mod=t["mod"], **gate)
    ops.linear(t["x"], W["up"], bias["up"], t["ffh"], epilogue=ops.EPI_BIAS_GELU)
    ops.linear(t["ffh"], W["down"], bias["down"], t["h"], mod=t["mod"], **gate)


meter = bench.EnergyMeter(0)


def measure(fn):
    fn(4)
    torch.cuda.synchronize()
    res = []
    for _ in range(REPS):
        j0 = meter.read()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn(LAYERS)
        e1.record()
        torch.cuda.synchronize()
        res.append((e0.elapsed_time(e1) / LAYERS, (meter.read() - j0) / LAYERS))
    return res


if mode == "serial":
    t = make(2)

    def run(layers):
        for _ in range(layers):
            pre_attn(t)
            attn(t)
            post_attn(t)
else:
    ta, tb = make(1), make(1)
    lo_pri, hi_pri = torch.cuda.Stream.priority_range()[0], torch.cuda.Stream.priority_range()[1]
    gemm_s = [torch.cuda.Stream(priority=hi_pri), torch.cuda.Stream(priority=hi_pri)]
    attn_s = [torch.cuda.Stream(priority=lo_pri), torch.cuda.Stream(priority=lo_pri)]

    def run(layers):
        main = torch.cuda.current_stream()
        for s in gemm_s + attn_s:
            s.wait_stream(main)
        seqs = [ta, tb]
        # sequence 1 starts half a layer late: its first GEMM phase waits for sequence 0's first attention to have started
        for l in range(layers):
            for i in (0, 1):
                with torch.cuda.stream(gemm_s[i]):
                    if l == 0 and i == 1:
                        gemm_s[1].wait_stream(gemm_s[0])
                    pre_attn(seqs[i])
                attn_s[i].wait_stream(gemm_s[i])
                with torch.cuda.stream(attn_s[i]):
                    attn(seqs[i])
                gemm_s[i].wait_stream(attn_s[i])
                with torch.cuda.stream(gemm_s[i]):
                    post_attn(seqs[i])
        for s in gemm_s + attn_s:
            main.wait_stream(s)

res = measure(run)
print(json.dumps({"mode": mode, "gemm_sms": os.environ.get("S2V_GEMM_SMS", "all"), "gemm_2cta": os.environ.get("S2V_GEMM_2CTA", "default"),
                  "ms_per_layer_both_sequences": [round(r[0], 3) for r in res], "joule_per_layer": [round(r[1], 2) for r in res],
                  "mean_power_w": [round(r[1] / r[0] * 1e3) for r in res]}), flush=True)
