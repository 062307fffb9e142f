"""Golden fixture for the T5 prompt encoder: a tiny random `transformers.T5EncoderModel` (gated-gelu, the
T5 v1.1 / CogVideoX text-encoder architecture) run by the INSTALLED transformers library itself — the reference imports that class
and nothing else (S/inference.py:13,185).  Run in the build container:  python tests/golden/make_golden_t5.py"""
import os

import torch
import transformers
from transformers import T5Config, T5EncoderModel

HERE = os.path.dirname(os.path.abspath(__file__))
torch.manual_seed(11)
cfg = dict(d_model=128, d_kv=64, d_ff=256, num_layers=2, num_heads=2, relative_attention_num_buckets=32, relative_attention_max_distance=128,
           feed_forward_proj="gated-gelu", vocab_size=96, dropout_rate=0.0)
m = T5EncoderModel(T5Config(**cfg)).eval()
with torch.no_grad():
    for n, p in m.named_parameters():
        if "layer_norm" in n:
            p.copy_(1.0 + 0.1 * torch.randn_like(p))
        elif "relative_attention_bias" in n:
            p.copy_(0.5 * torch.randn_like(p))
        elif n == "shared.weight":
            p.copy_(torch.randn_like(p))
        else:
            p.copy_(torch.randn_like(p) / p.shape[1] ** 0.5)
g = torch.Generator().manual_seed(12)
ids = torch.randint(0, 96, (2, 226), generator=g)
ids[:, 40:] = 0          # padded to 226 with token 0 like tokenizer(..., padding="max_length"); the reference passes NO mask
with torch.no_grad():
    out32 = m(ids)[0]
    m16 = T5EncoderModel(T5Config(**cfg)).eval()
    m16.load_state_dict(m.state_dict())
    m16 = m16.to(torch.bfloat16)          # what `text_encoder.to(device, dtype=weight_dtype)` leaves (S/inference.py:214): wo in bf16 too
    out16 = m16(ids)[0]
sd = {k: v.clone() for k, v in m.state_dict().items() if k != "encoder.embed_tokens.weight"}   # tied to shared.weight
torch.save(dict(cfg=cfg, state=sd, ids=ids, out_fp32=out32, out_bf16=out16.float(), transformers=transformers.__version__),
           os.path.join(HERE, "t5_tiny.pt"))
print("t5 golden ok", out32.shape, float(out32.abs().mean()), float((out16.float() - out32).norm() / out32.norm()))
