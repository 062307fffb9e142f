"""GPU parity of the T5 prompt encoder: the engine behind `attach_t5(transformers.T5EncoderModel)` against the
golden produced by the installed transformers library itself (tests/golden/t5_tiny.pt, make_golden_t5.py) and its kernels against
torch on the same operands.  Yardstick as everywhere: the library's own bf16 execution against its fp32 output."""
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
BF16 = torch.bfloat16


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import s2v_b200
    assert s2v_b200._lib.load().s2v_device_check(0) == 0, "not a B200"
    return torch.device("cuda:0")


def rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return float((a - b).norm() / (b.norm() + 1e-12))


def test_t5_kernels_vs_torch(dev, parity):
    import s2v_b200
    from s2v_b200 import _lib
    lib = _lib.load()
    st = torch.cuda.current_stream().cuda_stream
    g = torch.Generator().manual_seed(5)
    # RMSNorm with transformers' two roundings
    x = (3 * torch.randn(453, 4096, generator=g)).to(BF16).to(dev)
    w = (1 + 0.1 * torch.randn(4096, generator=g)).to(BF16).to(dev)
    out = torch.empty_like(x)
    _lib.check(lib.s2v_rmsnorm(x.data_ptr(), w.data_ptr(), out.data_ptr(), 453, 4096, 1e-6, st), "rmsnorm")
    var = x.float().pow(2).mean(-1, keepdim=True)
    want = w * (x * torch.rsqrt(var + 1e-6)).to(BF16)
    assert float((out.float() - want.float()).abs().max()) <= 2 ** -6 * float(want.float().abs().max())   # rsqrtf vs torch.rsqrt: 1 bf16 ulp
    parity.check("t5.rmsnorm", rel(out, want), default=2e-3, note="torch T5LayerNorm expression on the same bf16 input (CUDA)")
    # gated GELU
    gg = torch.randn(300, 2 * 1024, generator=g).to(BF16).to(dev)
    o = torch.empty(300, 1024, device=dev, dtype=BF16)
    _lib.check(lib.s2v_gated_gelu(gg.data_ptr(), o.data_ptr(), 300, 1024, st), "gated_gelu")
    want = F.gelu(gg[:, :1024].float(), approximate="tanh").to(BF16).float() * gg[:, 1024:].float()
    parity.check("t5.gated_gelu", rel(o, want), default=4e-3, note="fp32 gelu_new * gate on the same bf16 input")
    # embedding gather: exact
    table = torch.randn(100, 256, generator=g).to(BF16).to(dev)
    ids = torch.randint(0, 100, (37,), generator=g).to(dev)
    rows = torch.empty(37, 256, device=dev, dtype=BF16)
    _lib.check(lib.s2v_gather_rows(table.data_ptr(), ids.data_ptr(), rows.data_ptr(), 37, 256, 100, st), "gather")
    assert torch.equal(rows, table[ids])
    # attention with bias, no scaling.  S <= 256 runs the tensor-core kernel (score tiles per row: 8 for S <= 64, 16 for <= 128, 30 for <= 240,
    # 32 for <= 256 — every instantiation and its boundaries, ragged and odd sizes), longer sequences the scalar one (8 / 16 keys per lane)
    for (B, S, H) in [(2, 226, 3), (1, 300, 2), (1, 1, 1), (1, 33, 2), (1, 64, 1), (2, 65, 1), (1, 100, 2), (1, 128, 1), (1, 129, 1), (1, 240, 1),
                      (1, 241, 1), (1, 256, 2), (1, 257, 1), (1, 512, 1)]:
        qkv = torch.randn(B, S, 3 * H * 64, generator=g).mul(0.5).to(BF16).to(dev)
        bias = torch.randn(H, S, S, generator=g).to(BF16).to(dev)
        o = torch.full((B, S, H * 64), float("nan"), device=dev, dtype=BF16)
        _lib.check(lib.s2v_t5_attention(qkv.data_ptr(), bias.data_ptr(), o.data_ptr(), B, S, H, st), "t5_attention")
        q, k, v = [t.view(B, S, H, 64).transpose(1, 2).float() for t in qkv.chunk(3, dim=-1)]
        # transformers' rounding points: the score matmul and the bias add produce bf16 tensors, softmax runs in fp32 and is cast back
        sc = ((q @ k.transpose(2, 3)).to(BF16) + bias[None]).float()
        want = (torch.softmax(sc, dim=-1).to(BF16).float() @ v).transpose(1, 2).reshape(B, S, H * 64)
        assert torch.isfinite(o.float()).all()
        parity.check(f"t5.attention[B={B},S={S},H={H}]", rel(o, want), default=5e-3,
                     note="softmax(bf16(bf16(q k^T) + bias)) v with modeling_t5.py's rounding points, fp32 accumulation")
    assert lib.s2v_t5_attention(qkv.data_ptr(), bias.data_ptr(), o.data_ptr(), 1, 513, 1, st) == -2      # > 512 tokens refused
    assert lib.s2v_rmsnorm(None, None, None, 1, 8, 1e-6, st) == -1


def test_t5_encoder_vs_transformers_golden(dev, golden_dir, parity):
    """`attach_t5` on a real transformers.T5EncoderModel carrying the golden's weights in bf16 (what S/inference.py:214 leaves)."""
    transformers = pytest.importorskip("transformers")
    import s2v_b200
    fx = torch.load(os.path.join(golden_dir, "t5_tiny.pt"))
    m = transformers.T5EncoderModel(transformers.T5Config(**fx["cfg"])).eval()
    m.load_state_dict(fx["state"], strict=False)
    m = m.to(BF16).to(dev)
    s2v_b200.attach_t5(m)
    out = m(fx["ids"].to(dev))
    got = out[0]
    assert got is out.last_hidden_state and got.dtype == BF16 and got.shape == fx["out_fp32"].shape
    e_mine, e_ref = rel(got, fx["out_fp32"]), rel(fx["out_bf16"], fx["out_fp32"])
    parity.check("t5_encoder_tiny.vs_transformers_fp32", e_mine, ref=e_ref, default=max(1.5 * e_ref, 1e-2),
                 note="yardstick = transformers' own bf16 execution of the same model (the reference's run dtype)")
    with pytest.raises(RuntimeError, match="attention masks"):
        m(fx["ids"].to(dev), attention_mask=torch.zeros_like(fx["ids"]).to(dev))
    # the pipeline's prompt path runs on it unchanged (tokenizer stub: the real one is SentencePiece, a library)
    class Tok:
        def __call__(self, prompt, **kw):
            class R:
                input_ids = fx["ids"][: len(prompt)]
            return R()
    pipe = s2v_b200.CustomCogVideoXPipeline(Tok(), m, None, None, None)
    pe = pipe._get_t5_prompt_embeds(["a", "b"], 1, 226, dev, BF16)
    assert torch.equal(pe, got)


def test_t5_xxl_width_two_layers_vs_oracle(dev, parity):
    """Two layers at the real T5-XXL width (d_model 4096, 64 heads, d_ff 10240), 2 x 226 tokens, against the oracle on the host."""
    from oracle import t5_oracle as T
    from s2v_b200 import t5
    cfg = T.T5Config(num_layers=2, vocab_size=512)
    g = torch.Generator().manual_seed(3)
    sd = {"shared.weight": torch.randn(512, 4096, generator=g),
          "encoder.block.0.layer.0.SelfAttention.relative_attention_bias.weight": 0.5 * torch.randn(32, 64, generator=g),
          "encoder.final_layer_norm.weight": 1 + 0.1 * torch.randn(4096, generator=g)}
    for i in range(2):
        a, f = f"encoder.block.{i}.layer.0.", f"encoder.block.{i}.layer.1."
        for n in "qkvo":
            sd[a + f"SelfAttention.{n}.weight"] = torch.randn(4096, 4096, generator=g) / 64 * (0.25 if n in "qk" else 1.0)
        sd[a + "layer_norm.weight"] = 1 + 0.1 * torch.randn(4096, generator=g)
        sd[f + "layer_norm.weight"] = 1 + 0.1 * torch.randn(4096, generator=g)
        sd[f + "DenseReluDense.wi_0.weight"] = torch.randn(10240, 4096, generator=g) / 64
        sd[f + "DenseReluDense.wi_1.weight"] = torch.randn(10240, 4096, generator=g) / 64
        sd[f + "DenseReluDense.wo.weight"] = torch.randn(4096, 10240, generator=g) / 101
    p16 = {k: v.to(BF16) for k, v in sd.items()}
    ids = torch.randint(0, 512, (2, 226), generator=g)
    eng = t5.T5EncoderEngine({k: v.to(dev) for k, v in p16.items()}, 2, 64, 64)
    got = eng.forward(ids.to(dev))
    torch.cuda.synchronize()
    with torch.no_grad():
        exact = T.encoder_forward({k: v.float() for k, v in p16.items()}, cfg, ids)
        ref16 = T.encoder_forward(p16, cfg, ids)
    parity.check("t5_xxl_width_2_layers", rel(got, exact), ref=rel(ref16, exact), default=max(1.5 * rel(ref16, exact), 1e-2),
                 note="same bf16-rounded weights, fp32 oracle; yardstick = the oracle (= transformers' ops) executed in bf16")
