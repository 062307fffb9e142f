"""CPU-side checks (no GPU): the C-ABI library loads and exports every symbol the header declares, host logic
(scheduler tables, RoPE / sincos tables, LoRA layout + packing, sharding plans) matches the reference goldens, and the
product path FAILS LOUDLY without a B200 (no CPU / eager fallback)."""
import os
import re

import numpy as np
import pytest
import torch

import s2v_b200
from s2v_b200 import _lib, engine, lora, modules, ops, parallel, scheduler, tables

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ------------------------------------------------------------------ C ABI
def _header_symbols():
    src = open(os.path.join(ROOT, "include", "s2v_b200.h")).read()
    return sorted(set(re.findall(r"S2V_API\s+(?:const\s+char\*|int64_t|int)\s+(s2v_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    syms = _header_symbols()
    assert len(syms) >= 19
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/s2v_b200.h but not exported"
    # and the Python binding knows every compute entry point
    for s in syms:
        if s not in ("s2v_last_error", "s2v_workspace_bytes"):     # the two entries that do not return an int status
            assert s in _lib.SIGNATURES, s
    assert "s2v_workspace_bytes" in syms
    assert lib.s2v_abi_version() == 1


def test_workspace_bytes_query():
    """SURVEY §8b `s2v_workspace_bytes`: host-only arithmetic, usable without a GPU; the cfg-3 geometry needs ~2.6 GB."""
    lib = _lib.load()
    B, S, D = 2, 19126, 3072
    got = lib.s2v_workspace_bytes(B, S, D, 4 * D, 3 * 128, 85)
    up = lambda b: (b + 255) // 256 * 256  # noqa: E731
    rows = B * S
    want = 3 * up(rows * D * 2) + up(rows * 3 * D * 2) + up(rows * 4 * D * 2) + up(rows * 384 * 2) + up(85 * B * 6 * D * 4)
    assert got == want and 2.3e9 < got < 2.8e9
    assert lib.s2v_workspace_bytes(0, S, D, 4 * D, 0, 1) < 0 and b"bad geometry" in lib.s2v_last_error()


def test_linear_args_struct_matches_header_layout():
    # field order and sizes of s2v_linear_args (LP64): 8-byte pointers / int64, 4-byte ints and float
    import ctypes as C
    assert C.sizeof(_lib.LinearArgs) == 8 * 9 + 4 * 2 + 8 * 2 + 4 * 4 + 4 + 4 + 8 + 4 * 5 + 4  # incl. tail padding
    assert _lib.LinearArgs.mod.offset % 8 == 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    lib = _lib.load()
    assert lib.s2v_device_check(0) == -3  # S2V_E_NO_DEVICE
    x = torch.zeros(8, 64, dtype=torch.bfloat16)
    with pytest.raises(RuntimeError):
        ops.linear(x, x, None, torch.zeros(8, 8, dtype=torch.bfloat16))
    m = modules.CogVideoXTransformer3DModel(num_attention_heads=1, num_layers=1, time_embed_dim=64, text_embed_dim=64).to(torch.bfloat16)
    with pytest.raises(RuntimeError):
        m(torch.zeros(2, 1, 16, 4, 4), torch.zeros(1, 1, 16, 4, 4), torch.zeros(2, 4, 64), torch.tensor([1, 1]), eval=True)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "disentangled-subject-to-vid_b200")
    for f in os.listdir(pkg):
        if f.endswith(".py"):
            src = open(os.path.join(pkg, f)).read()
            assert not re.search(r"^\s*(from|import)\s+\.*oracle", src, flags=re.M), f
            assert "import_module(\"oracle" not in src and "s2v_oracle" not in src, f


# ------------------------------------------------------------------ scheduler host arithmetic (bit-exact vs reference)
@pytest.mark.parametrize("tag,snr", [("5b", 1.0), ("2b", 3.0)])
def test_scheduler_tables_bit_exact(golden_dir, tag, snr):
    g = np.load(os.path.join(golden_dir, "scheduler_ddim.npz"))
    s = scheduler.CogVideoXDDIMScheduler.for_cogvideox(snr)
    assert np.array_equal(s.alphas_cumprod.numpy(), g[f"alphas_cumprod_{tag}"])
    for n in (50, 7, 30):
        s.set_timesteps(n)
        assert np.array_equal(s.timesteps.numpy(), g[f"timesteps_{tag}_{n}"])
        assert s._timesteps_host == [int(t) for t in g[f"timesteps_{tag}_{n}"]]


def test_scheduler_errors_match_reference():
    s = scheduler.CogVideoXDDIMScheduler.for_cogvideox(1.0)
    with pytest.raises(ValueError, match="set_timesteps"):
        s.coefficients(999)
    with pytest.raises(ValueError, match="cannot be larger"):
        s.set_timesteps(1001)


@pytest.mark.parametrize("tag,snr", [("5b", 1.0), ("2b", 3.0)])
def test_scheduler_coefficients_reproduce_cpu_reference_trace(golden_dir, tag, snr):
    """With scalar_semantics='cpu' the four fp32 coefficients + the kernel's rounding model (emulated here with torch
    elementwise ops) reproduce the reference scheduler's 50-step trace bit for bit."""
    g = np.load(os.path.join(golden_dir, "scheduler_ddim.npz"))
    s = scheduler.CogVideoXDDIMScheduler.for_cogvideox(snr)
    s.scalar_semantics = "cpu"
    s.set_timesteps(50)
    x = torch.from_numpy(g[f"trace_{tag}_sample0"]).to(torch.bfloat16)
    rb = lambda z: z.to(torch.bfloat16).float()  # noqa: E731
    for i, t in enumerate(s._timesteps_host):
        sa, sb, a, b = (torch.tensor(c, dtype=torch.float32) for c in s.coefficients(t))
        v = torch.from_numpy(g[f"trace_{tag}_model_out"][i])
        x0 = rb(sa * x.float()) - sb * v
        prev = rb(a * x.float()) + b * x0
        assert np.array_equal(prev.numpy(), g[f"trace_{tag}_prev"][i]), i
        x = prev.to(torch.bfloat16)


# ------------------------------------------------------------------ tables
def test_rope_tables_bit_exact(golden_dir):
    import hashlib
    g = np.load(os.path.join(golden_dir, "rope.npz"))
    cos, sin = tables.rope_table_3d(64, ((0, 0), (8, 8)), (8, 8), 2)
    assert np.array_equal(cos.numpy(), g["small_cos"]) and np.array_equal(sin.numpy(), g["small_sin"])
    for tag, h, w, T in (("480x720_T14", 480, 720, 14), ("720x1280_T14", 720, 1280, 14), ("480x720_T3", 480, 720, 3)):
        cos, sin = tables.joint_rope_table(h, w, T - 1)
        assert hashlib.sha256(cos.numpy().tobytes()).hexdigest() == str(g[f"{tag}_cos_sha"])
        assert hashlib.sha256(sin.numpy().tobytes()).hexdigest() == str(g[f"{tag}_sin_sha"])
    for src in ((30, 45), (45, 80), (60, 60), (8, 12)):
        (a, b), (c, d) = tables.crop_region_for_grid(src, 45, 30)
        assert [a, b, c, d] == list(g[f"crop_{src[0]}x{src[1]}"])


def test_sincos_table_matches_oracle():
    from oracle import s2v_oracle as O
    t = tables.sincos_table_3d(1920, 6, 4, 3, 1.875, 1.0)
    ref = torch.from_numpy(O.sincos_pos_embed_3d(1920, 6, 4, 3, 1.875, 1.0)).flatten(0, 1).float()
    assert torch.equal(t, ref)


# ------------------------------------------------------------------ LoRA layout + packing
def _tiny_model(**kw):
    return modules.CogVideoXTransformer3DModel(num_attention_heads=2, num_layers=2, time_embed_dim=64, text_embed_dim=64, **kw)


def test_state_dict_keys_match_reference_schema():
    from oracle import s2v_oracle as O
    m = _tiny_model(use_rotary_positional_embeddings=True)
    cfg = O.TransformerConfig(num_attention_heads=2, num_layers=2, time_embed_dim=64, text_embed_dim=64)
    want = O.param_shapes(cfg)
    got = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert got == {k: tuple(v) for k, v in want.items()}


def test_lora_injection_targets_and_packing():
    m = _tiny_model().to(torch.bfloat16)
    names = lora.inject_lora(m, r=8, alpha=4.0)
    per_block = {"attn1.to_q", "attn1.to_k", "attn1.to_v", "attn1.to_out.0", "norm1.linear", "norm2.linear", "ff.net.0.proj", "ff.net.2"}
    want = {f"transformer_blocks.{i}.{s}" for i in range(2) for s in per_block} | {"patch_embed.proj", "patch_embed.text_proj"}
    assert set(names) == want  # proj_out, norm_out.linear, time_embedding.* are NOT matched
    sd = {f"transformer.{n}.lora_{ab}.weight": torch.randn_like(getattr(m.get_submodule(n), f"lora_{ab}")["default"].weight)
          for n in names for ab in "AB"}
    assert lora.load_lora_state_dict(m, sd) == 2 * len(names)
    # the PEFT key form written by convert_unet_state_dict_to_peft (S/inference.py:94) loads too ...
    peft_form = {k.replace(".weight", ".default.weight")[len("transformer."):]: v + 1 for k, v in sd.items()}
    assert lora.load_lora_state_dict(m, peft_form) == 2 * len(names)
    k0 = f"{names[0]}.lora_B.default.weight"
    assert torch.equal(m.get_submodule(names[0]).lora_B["default"].weight, peft_form[k0])
    # ... and a checkpoint that would load nothing, half of the factors, or foreign keys is an error, not a silent B = 0 run
    with pytest.raises(KeyError):
        lora.load_lora_state_dict(m, {})
    with pytest.raises(KeyError):
        lora.load_lora_state_dict(m, {k: v for k, v in sd.items() if "lora_B" in k})
    with pytest.raises(KeyError):
        lora.load_lora_state_dict(m, dict(sd, **{"transformer.proj_out.weight": torch.zeros(1)}))
    assert lora.load_lora_state_dict(m, dict(sd, **{"transformer.proj_out.weight": torch.zeros(1)}), strict=False) == 2 * len(names)
    pb = engine.pack_block(m.transformer_blocks[0])
    D = 128
    assert pb.qkv.w.shape == (3 * D, D) and pb.qkv.a.shape == (24, D) and pb.qkv.bb.shape == (3 * D, 8) and pb.qkv.group_n == D
    assert pb.qkv.scale == 0.5 and pb.ff1.bb.shape == (4 * D, 8) and pb.norm1.a.shape == (8, 64)
    at = m.transformer_blocks[0].attn1
    assert torch.equal(pb.qkv.w[D:2 * D], at.to_k.base_layer.weight) and torch.equal(pb.qkv.bb[2 * D:], at.to_v.lora_B["default"].weight)
    # merged mode folds s*B*A into W with one rounding
    pm = engine.pack_block(m.transformer_blocks[0], merge_lora=True)
    w = at.to_q.base_layer.weight.float() + 0.5 * at.to_q.lora_B["default"].weight.float() @ at.to_q.lora_A["default"].weight.float()
    assert pm.qkv.a is None and torch.equal(pm.qkv.w[:D], w.to(torch.bfloat16))


def test_engine_dtype_policy():
    """fp32 parameters are refused; fp16 ones (how S/inference.py:191,210 loads CogVideoX-2B) are packed as bf16 copies."""
    m = _tiny_model()
    with pytest.raises(RuntimeError, match="bfloat16"):
        engine.pack_block(m.transformer_blocks[0])
    h = _tiny_model().half()
    with pytest.warns(UserWarning, match="float16"):
        engine._FP16_WARNED = False
        pb = engine.pack_block(h.transformer_blocks[0])
    assert pb.qkv.w.dtype == torch.bfloat16 and pb.ln1_w.dtype == torch.bfloat16
    assert torch.equal(pb.out.w, h.transformer_blocks[0].attn1.to_out[0].weight.to(torch.bfloat16))


def test_packed_snapshots_are_dropped_when_weights_change():
    """the engine's q|k|v / LoRA concatenations are copies; load_state_dict, .to(), inject_lora and load_lora_state_dict
    must drop them (here on CPU: the caches are plain attributes, no kernel runs)."""
    m = _tiny_model().to(torch.bfloat16)
    blk = m.transformer_blocks[0]
    blk._pb = engine.pack_block(blk)
    blk.attn1._pb = object()
    m._engine = object()
    m.load_state_dict(m.state_dict())
    assert m._engine is None and blk._pb is None and blk.attn1._pb is None
    blk._pb, m._engine = object(), object()
    lora.inject_lora(m, r=8, alpha=4.0)
    assert m._engine is None and blk._pb is None
    blk._pb, m._engine = object(), object()
    m.to(torch.bfloat16)
    assert m._engine is None and blk._pb is None
    # a stock model bound with attach() is re-packed through the same paths
    calls = []

    class FakeEngine:
        def repack(self):
            calls.append(1)

    m._s2v_engine = FakeEngine()
    lora.invalidate_packed(m)
    assert calls == [1]


# ------------------------------------------------------------------ pipeline input validation (reference error surface)
def test_pipeline_errors_match_reference():
    from s2v_b200 import CustomCogVideoXPipeline
    m = _tiny_model().to(torch.bfloat16)
    pipe = CustomCogVideoXPipeline(None, None, m, None, scheduler.CogVideoXDDIMScheduler.for_cogvideox())
    with pytest.raises(ValueError, match="less than or equal to 49"):
        pipe(prompt_embeds=torch.zeros(1, 4, 64), negative_prompt_embeds=torch.zeros(1, 4, 64), num_frames=53)
    with pytest.raises(ValueError, match="divisible by 8"):
        pipe(prompt_embeds=torch.zeros(1, 4, 64), negative_prompt_embeds=torch.zeros(1, 4, 64), height=481)
    with pytest.raises(ValueError, match="Provide either"):
        pipe()
    with pytest.raises(ValueError, match="Cannot forward both"):
        pipe(prompt="a", prompt_embeds=torch.zeros(1, 4, 64))


# ------------------------------------------------------------------ sharding plans
def test_shard_plans():
    assert parallel.plan(8, 8, 3).prompts == [3] and parallel.plan(8, 8, 3).mode == "prompt"
    assert parallel.plan(8, 4, 1).prompts == [1, 5]
    p = parallel.plan(1, 2, 1)
    assert p.mode == "cfg" and p.cfg_half == 1 and p.pair_ranks == [0, 1] and p.prompts == [0]
    assert parallel.plan(2, 4, 2).prompts == [1] and parallel.plan(2, 4, 3).cfg_half == 1
    with pytest.raises(ValueError):
        parallel.plan(3, 2, 0)
    pe = torch.arange(8).view(8, 1, 1).float()  # [neg0..3, pos0..3]
    assert parallel.select_prompt_embeds(pe, 4, parallel.plan(4, 2, 1)).flatten().tolist() == [1, 3, 5, 7]
    assert parallel.select_prompt_embeds(pe, 4, parallel.plan(4, 8, 5)).flatten().tolist() == [6]


def test_vae_state_dict_halves_load_strictly_and_full_checkpoint_schema():
    """AutoencoderKLCogVideoX keeps the reference's state-dict schema for BOTH halves; a decoder-only (or encoder-only) dict
    loads that half strictly, a dict with a hole raises (autoencoder_kl_cogvideox.py:1020-1115 parameter tree)."""
    import s2v_b200
    from oracle import vae_oracle as V
    cfg = V.VaeConfig(block_out_channels=(64, 64, 64, 64), layers_per_block=1, sample_height=64, sample_width=96)
    m = s2v_b200.AutoencoderKLCogVideoX(block_out_channels=cfg.block_out_channels, layers_per_block=1, sample_height=64, sample_width=96)
    dec, enc = V.synth_decoder_params(cfg, 1), V.synth_encoder_params(cfg, 2)
    assert set(m.state_dict()) == set(dec) | set(enc)
    r = m.load_state_dict(dec)
    assert all(k.startswith("encoder.") for k in r.missing_keys) and not r.unexpected_keys
    r = m.load_state_dict(enc)
    assert all(k.startswith("decoder.") for k in r.missing_keys) and not r.unexpected_keys
    assert not m.load_state_dict({**dec, **enc}).missing_keys
    holed = dict(dec)
    holed.pop("decoder.conv_in.conv.weight")
    with pytest.raises(RuntimeError):
        m.load_state_dict(holed)
    with pytest.raises(RuntimeError):
        m.load_state_dict({**dec, "decoder.bogus.weight": dec["decoder.conv_in.conv.bias"]})


def test_diagonal_gaussian_and_reference_image_input_checks():
    """DiagonalGaussianDistribution follows D/models/autoencoders/vae.py:767-820 (chunk, clamp, std, sample = mean + std * noise with
    the generator's draw, mode = mean); encode_reference_image refuses anything but an RGB uint8 image before touching the device."""
    import numpy as np
    import s2v_b200
    from oracle import vae_oracle as V
    from s2v_b200.vae import DiagonalGaussianDistribution
    g = torch.Generator().manual_seed(5)
    mom = torch.randn(2, 8, 1, 3, 4, generator=g) * 3
    mom[0, 4, 0, 0, 0], mom[0, 5, 0, 0, 0] = 50.0, -50.0            # logvar beyond the clamp on both sides
    d = DiagonalGaussianDistribution(mom)
    assert torch.equal(d.mode(), mom[:, :4]) and float(d.logvar.max()) == 20.0 and float(d.logvar.min()) == -30.0
    noise = torch.randn(d.mean.shape, generator=torch.Generator().manual_seed(9), dtype=mom.dtype)
    assert torch.equal(d.sample(torch.Generator().manual_seed(9)), V.gaussian_sample(mom, noise))
    assert torch.equal(DiagonalGaussianDistribution(mom, deterministic=True).sample(torch.Generator().manual_seed(1)), mom[:, :4])

    class _Vae(torch.nn.Module):   # never reached: the input checks come first
        def __init__(self):
            super().__init__()
            self.p = torch.nn.Parameter(torch.zeros(1))
    for bad in (np.zeros((4, 4, 3), dtype=np.float32), np.zeros((4, 4), dtype=np.uint8), np.zeros((4, 4, 4), dtype=np.uint8)):
        with pytest.raises(ValueError):
            s2v_b200.encode_reference_image(_Vae(), bad)
