"""T5-XXL prompt encode (24 layers, d_model 4096, 64 heads, d_ff 10240; 2 prompts x 226 tokens, random weights) on one B200:
transformers' own eager module against the attached B200 engine, CUDA events, medians."""
import json
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import transformers

import s2v_b200

dev = torch.device("cuda:0")
cfg = transformers.T5Config(d_model=4096, d_kv=64, d_ff=10240, num_layers=24, num_heads=64, relative_attention_num_buckets=32,
                            relative_attention_max_distance=128, feed_forward_proj="gated-gelu", vocab_size=32128, dropout_rate=0.0)
with torch.device("meta"):
    m = transformers.T5EncoderModel(cfg)
m = m.to_empty(device=dev).eval()
g = torch.Generator(device=dev).manual_seed(0)
with torch.no_grad():
    for n, p in m.named_parameters():
        if "layer_norm" in n:
            p.copy_(1 + 0.1 * torch.randn(p.shape, device=dev, generator=g))
        else:
            p.copy_(torch.randn(p.shape, device=dev, generator=g) / p.shape[-1] ** 0.5 * (0.1 if (".q." in n or ".k." in n) else 1.0))
    m.encoder.embed_tokens.weight = m.shared.weight        # to_empty() breaks the tie; a loaded checkpoint has it
ids = torch.randint(0, 32128, (2, 226), device=dev, generator=g)


def timed(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ms = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = fn(); e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    return statistics.median(ms), out


with torch.no_grad():
    m32 = m.float()
    out32 = m32(ids)[0]                      # fp32 execution of the same weights: the yardstick for both bf16 executions
    m = m32.to(torch.bfloat16)
    hf_ms, hf_out = timed(lambda: m(ids)[0])
    s2v_b200.attach_t5(m)
    my_ms, my_out = timed(lambda: m(ids)[0])
params = sum(p.numel() for p in m.parameters())
print(json.dumps({"model": "T5 v1.1 XXL encoder, random weights, bf16", "tokens": [2, 226], "params_G": round(params / 1e9, 2),
                  "transformers_eager_ms": round(hf_ms, 2), "s2v_engine_ms": round(my_ms, 2), "speedup": round(hf_ms / my_ms, 2),
                  "weight_read_floor_ms": round(params * 2 / 6549.8e9 * 1e3, 2),
                  "rel_err_vs_fp32_execution": {"s2v_engine": float((my_out.float() - out32).norm() / out32.norm()),
                                                "transformers_bf16": float((hf_out.float() - out32).norm() / out32.norm())}}))
