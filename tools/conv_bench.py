"""Time the implicit-GEMM convolution at VAE shapes (CUDA events), with and without the residual input.
    S2V_CONV_T=0|1 python tools/conv_bench.py"""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from s2v_b200 import _lib

dev = "cuda"
lib = _lib.load()
torch.manual_seed(0)


def run(T, Hh, W, cin, cout, taps, with_res):
    x = torch.randn(T + 2, Hh + 2, W + 2, cin, device=dev).to(torch.bfloat16)
    w = (torch.randn(cout, taps * cin, device=dev) / (taps * cin) ** 0.5).to(torch.bfloat16)
    b = torch.zeros(cout, device=dev, dtype=torch.bfloat16)
    o = torch.empty(T + 2, Hh + 2, W + 2, cout, device=dev, dtype=torch.bfloat16)
    r = torch.randn(T + 2, Hh + 2, W + 2, cout, device=dev).to(torch.bfloat16) if with_res else None
    a = _lib.ConvArgs()
    a.x, a.ldx, a.w, a.ldw, a.bias, a.out, a.ldo = x.data_ptr(), cin, w.data_ptr(), taps * cin, b.data_ptr(), o.data_ptr(), cout
    a.res, a.ldres = (r.data_ptr(), cout) if with_res else (None, 0)
    a.T, a.t_pad, a.Hp, a.Wp, a.cin, a.cout, a.taps = T, 2, Hh + 2, W + 2, cin, cout, taps
    st = torch.cuda.current_stream().cuda_stream
    f = lambda: _lib.check(lib.s2v_conv_gemm(C.byref(a), st), "conv")  # noqa: E731
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        f()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    fl = 2.0 * T * (Hh + 2) * (W + 2) * cout * taps * cin
    print(json.dumps({"conv_t": os.environ.get("S2V_CONV_T", "1"), "T": T, "H": Hh, "W": W, "cin": cin, "cout": cout, "taps": taps, "res": with_res,
                      "ms": round(ms, 4), "tflops": round(fl / ms / 1e9, 1)}), flush=True)


for shape in [(9, 240, 360, 128, 128, 27), (9, 240, 360, 256, 128, 27), (8, 240, 360, 128, 128, 27), (9, 80, 144, 128, 128, 27), (3, 240, 360, 128, 128, 27)]:
    for res in (False, True):
        run(*shape, res)
