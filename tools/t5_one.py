import os, sys
sys.path.insert(0, os.getcwd())
os.environ["S2V_T5_GRAPH"] = "0"
import torch, transformers
import s2v_b200
dev = torch.device("cuda:0")
cfg = transformers.T5Config(d_model=4096, d_kv=64, d_ff=10240, num_layers=24, num_heads=64, relative_attention_num_buckets=32,
                            relative_attention_max_distance=128, feed_forward_proj="gated-gelu", vocab_size=32128, dropout_rate=0.0)
with torch.device("meta"):
    m = transformers.T5EncoderModel(cfg)
m = m.to_empty(device=dev).eval().to(torch.bfloat16)
with torch.no_grad():
    for n, p in m.named_parameters():
        p.copy_(torch.randn(p.shape, device=dev) * 0.02 + (1.0 if "layer_norm" in n else 0.0))
ids = torch.randint(0, 32128, (2, 226), device=dev)
s2v_b200.attach_t5(m)
with torch.no_grad():
    for _ in range(3):
        m(ids)[0]
torch.cuda.synchronize()
print("done")
