"""Energy per launch of the big kernels: each arm runs back to back for ~3 s while `nvidia-smi` samples board power and SM clock
every 100 ms; J/launch = mean power x mean launch time.  Arms: torch SDPA (cuDNN) and this repo's attention variants at the cfg-3
shape, the FFN-up projection, and (for reference) an idle interval.  Evidence for the "the step is energy-bound" analysis in
profiles/r02_summary.md — under the 1 kW cap a kernel's sustained time IS its energy."""
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.nn.functional as F

from s2v_b200 import ops

dev = torch.device("cuda:0")
B, S, H = 2, 19126, 48
torch.manual_seed(0)
qkv = torch.randn(B, S, 3 * H * 64, device=dev).to(torch.bfloat16)
out = torch.empty(B, S, H * 64, device=dev, dtype=torch.bfloat16)
q4 = qkv.view(B, S, 3, H, 64)
qs, ks, vs = (q4[:, :, i].transpose(1, 2) for i in range(3))
exp = C.CDLL(os.path.join(ROOT, "tools", "bin", "libattn_exp.so"))
exp.s2v_attn_fwd_exp.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_int32, C.c_int32, C.c_int32,
                                 C.c_void_p, C.c_void_p]
M, D = B * S, 3072
x = torch.randn(M, D, device=dev).to(torch.bfloat16)
w = (0.02 * torch.randn(4 * D, D, device=dev)).to(torch.bfloat16)
bias = torch.zeros(4 * D, device=dev, dtype=torch.bfloat16)
y = torch.empty(M, 4 * D, device=dev, dtype=torch.bfloat16)


def exp_arm(variant):
    return lambda: exp.s2v_attn_fwd_exp(qkv.data_ptr(), out.data_ptr(), B, S, H, 0.125, variant, 1, 200, None, torch.cuda.current_stream().cuda_stream)


arms = {"sdpa_cudnn": lambda: F.scaled_dot_product_attention(qs, ks, vs), "attn_shipped": lambda: ops.attention(qkv, out, H),
        "attn_r1_lo": exp_arm(0), "attn_hi_mc": exp_arm(3), "attn_bk80_hi_mc": exp_arm(7),
        "gemm_ffn_up": lambda: ops.linear(x, w, bias, y, epilogue=ops.EPI_BIAS_GELU),
        "torch_matmul_ffn_up": lambda: torch.matmul(x, w.t(), out=y)}
only = sys.argv[1:] or list(arms)


class Sampler:
    def __init__(self):
        self.rows = []
        self.p = subprocess.Popen(["nvidia-smi", "-i", "0", "--query-gpu=power.draw,clocks.sm,temperature.gpu", "--format=csv,noheader,nounits",
                                   "-lms", "100"], stdout=subprocess.PIPE, text=True)
        threading.Thread(target=self._pump, daemon=True).start()

    def _pump(self):
        for line in self.p.stdout:
            try:
                self.rows.append((time.time(),) + tuple(float(c) for c in line.split(",")))
            except ValueError:
                pass

    def between(self, t0, t1):
        r = [x for x in self.rows if t0 <= x[0] <= t1]
        n = max(len(r), 1)
        return sum(x[1] for x in r) / n, sum(x[2] for x in r) / n, sum(x[3] for x in r) / n, len(r)


smp = Sampler()
time.sleep(1.5)
t0 = time.time(); time.sleep(1.0)
print(json.dumps({"arm": "idle", "power_w": round(smp.between(t0, time.time())[0], 1)}), flush=True)
dur = float(os.environ.get("SECONDS", "3"))
for name in only:
    fn = arms[name]
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    n = 0
    t_start = time.time()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    while time.time() - t_start < dur:
        for _ in range(10):
            fn()
        n += 10
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    t_end = time.time()
    ms = e0.elapsed_time(e1) / n
    pw, mhz, temp, k = smp.between(t_start + 0.7, t_end)   # skip the ramp
    print(json.dumps({"arm": name, "ms_per_launch": round(ms, 3), "power_w": round(pw, 1), "sm_mhz": round(mhz), "temp_c": round(temp),
                      "joule_per_launch": round(pw * ms / 1e3, 3), "samples": k}), flush=True)
    time.sleep(1.0)
smp.p.terminate()
