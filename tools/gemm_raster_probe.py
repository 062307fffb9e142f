"""The four big projections of a cfg-3 block (M = 2 x 19126 rows) launched a few times each, with CUDA-event timing: the target
of `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum` when comparing rasterisation group heights
(S2V_GEMM_GROUP_M=16|32|48|64 python tools/gemm_raster_probe.py)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from s2v_b200 import ops

dev = torch.device("cuda:0")
torch.manual_seed(0)
M, D, r = 2 * 19126, 3072, 128
bf = torch.bfloat16
x = torch.randn(M, D, device=dev).to(bf)
x4 = torch.randn(M, 4 * D, device=dev).to(bf)
mod = torch.randn(2, 6 * D, device=dev)
shapes = {
    "qkv": (x, 3 * D, dict()),
    "out": (x, D, dict(epilogue=ops.EPI_GATE_RESIDUAL, mod=mod, gate_off_text=5 * D, gate_off_other=2 * D, rows_per_batch=19126, text_len=226)),
    "ffn_up": (x, 4 * D, dict(epilogue=ops.EPI_BIAS_GELU)),
    "ffn_down": (x4, D, dict(epilogue=ops.EPI_GATE_RESIDUAL, mod=mod, gate_off_text=5 * D, gate_off_other=2 * D, rows_per_batch=19126, text_len=226)),
}
only = sys.argv[1:] or list(shapes)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for name in only:
    inp, N, kw = shapes[name]
    K = inp.shape[1]
    w = (0.02 * torch.randn(N, K, device=dev)).to(bf)
    b = (0.02 * torch.randn(N, device=dev)).to(bf)
    groups = 3 if name == "qkv" else 1
    a = (0.02 * torch.randn(groups * r, K, device=dev)).to(bf)
    bb = (0.02 * torch.randn(N, r, device=dev)).to(bf)
    t = torch.empty(M, groups * r, device=dev, dtype=bf)
    ops.linear(inp, a, None, t, alpha=0.5)
    out = torch.zeros(M, N, device=dev, dtype=bf)
    ms = []
    for i in range(int(os.environ.get("ITERS", "5"))):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.linear(inp, w, b, out, lora_t=t, lora_b=bb, lora_group_n=N // groups if groups > 1 else 0, **kw)
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    fl = 2.0 * M * N * (K + r)
    best = min(ms[1:])
    print(json.dumps({"gemm": name, "group_m": os.environ.get("S2V_GEMM_GROUP_M", "default"), "M": M, "N": N, "K": K, "ms": round(best, 4),
                      "tflops": round(fl / best / 1e9, 1), "algorithmic_read_MB": round((M * K + N * K + M * groups * r + N * r) * 2 / 1e6, 1),
                      "algorithmic_write_MB": round(M * N * 2 / 1e6, 1)}), flush=True)
    del w, out
