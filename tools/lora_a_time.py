import sys, os, json, statistics
sys.path.insert(0, os.getcwd())
import torch
from s2v_b200 import ops
dev="cuda"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timed(fn, n=15):
    for _ in range(3): fn()
    torch.cuda.synchronize(); ms=[]
    for _ in range(n):
        flush.zero_()
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ms.append(e0.elapsed_time(e1))
    return statistics.median(ms)
for (M,N,K) in [(38252,128,3072),(37888,128,3072),(38252,384,3072),(37888,384,3072),(38252,128,12288),(37888,128,12288),(18944,128,3072),(19126,128,3072)]:
    x = torch.randn(M, K, device=dev).to(torch.bfloat16)
    w = (0.02 * torch.randn(N, K, device=dev)).to(torch.bfloat16)
    o = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    t = timed(lambda: ops.linear(x, w, None, o))
    print(json.dumps({"M":M,"N":N,"K":K,"ms":round(t,4),"x_GB_per_s":round(M*K*2/t/1e6,1)}), flush=True)
