"""Pin the CPU oracle (oracle/s2v_oracle.py) against fixtures produced by the reference's own modules
(tests/golden/make_golden.py).  Integer / fp64 scheduler + RoPE-index arithmetic: bit-exact.  fp32 block / transformer /
loop: tight tolerance (same torch ops, but op ORDER differs slightly, e.g. three LayerNorm calls vs one)."""
import hashlib
import os

import numpy as np
import pytest
import torch

from oracle import s2v_oracle as O


def _sha(t):
    return hashlib.sha256(t.detach().contiguous().numpy().tobytes()).hexdigest()


def _cfg(d):
    return O.TransformerConfig(**d)


# ---------------------------------------------------------------- scheduler (bit-exact)
@pytest.mark.parametrize("tag,snr", [("5b", 1.0), ("2b", 3.0)])
def test_alphas_cumprod_bit_exact(golden_dir, tag, snr):
    g = np.load(os.path.join(golden_dir, "scheduler_ddim.npz"))
    ac = O.ddim_alphas_cumprod(snr_shift_scale=snr)
    assert ac.dtype == np.float64
    assert np.array_equal(ac, g[f"alphas_cumprod_{tag}"])
    # known-answer scalars recorded in SURVEY.md §8a row S / BASELINE.md §5
    assert ac[999] == 0.0
    assert ac[979] == (8.578784340054837e-05 if tag == "5b" else 2.7194360400217652e-05)


@pytest.mark.parametrize("n", [50, 7, 30])
def test_timesteps_bit_exact(golden_dir, n):
    g = np.load(os.path.join(golden_dir, "scheduler_ddim.npz"))
    ts = O.ddim_timesteps(n)
    assert np.array_equal(ts, g[f"timesteps_5b_{n}"])
    if n == 50:
        assert ts[0] == 999 and ts[1] == 979 and ts[-1] == 19


@pytest.mark.parametrize("tag,snr", [("5b", 1.0), ("2b", 3.0)])
def test_ddim_step_trace_bit_exact(golden_dir, tag, snr):
    g = np.load(os.path.join(golden_dir, "scheduler_ddim.npz"))
    ac = O.ddim_alphas_cumprod(snr_shift_scale=snr)
    sample = torch.from_numpy(g[f"trace_{tag}_sample0"]).to(torch.bfloat16)
    for i, t in enumerate(O.ddim_timesteps(50)):
        mo = torch.from_numpy(g[f"trace_{tag}_model_out"][i])
        prev, x0 = O.ddim_step(ac, mo, int(t), sample, 50)
        assert prev.dtype == torch.float32
        assert np.array_equal(prev.numpy(), g[f"trace_{tag}_prev"][i]), f"step {i}"
        assert np.array_equal(x0.numpy(), g[f"trace_{tag}_x0"][i]), f"step {i}"
        sample = prev.to(torch.bfloat16)


# ---------------------------------------------------------------- RoPE (bit-exact)
def test_rope_small_bit_exact(golden_dir):
    g = np.load(os.path.join(golden_dir, "rope.npz"))
    cos, sin = O.rope_3d_tables(64, ((0, 0), (8, 8)), (8, 8), 2)
    assert np.array_equal(cos.numpy(), g["small_cos"]) and np.array_equal(sin.numpy(), g["small_sin"])


@pytest.mark.parametrize("tag,h,w,T", [("480x720_T14", 480, 720, 14), ("720x1280_T14", 720, 1280, 14), ("480x720_T3", 480, 720, 3)])
def test_rope_pipeline_tables_bit_exact(golden_dir, tag, h, w, T):
    g = np.load(os.path.join(golden_dir, "rope.npz"))
    (vc, vs), (rc, rs) = O.pipeline_rope_tables(h, w, T - 1)
    cos, sin = torch.cat([rc, vc]), torch.cat([rs, vs])
    assert list(cos.shape) == list(g[f"{tag}_shape"])
    assert _sha(cos) == str(g[f"{tag}_cos_sha"]) and _sha(sin) == str(g[f"{tag}_sin_sha"])
    rows = g[f"{tag}_rows"]
    assert np.array_equal(cos[rows].numpy(), g[f"{tag}_cos_rows"])


def test_crop_region(golden_dir):
    g = np.load(os.path.join(golden_dir, "rope.npz"))
    for src in ((30, 45), (45, 80), (60, 60), (8, 12)):
        (a, b), (c, d) = O.resize_crop_region_for_grid(src, 45, 30)
        assert [a, b, c, d] == list(g[f"crop_{src[0]}x{src[1]}"])


# ---------------------------------------------------------------- block / transformer / loop (fp32)
def _rope_small():
    cos, sin = O.rope_3d_tables(64, ((0, 0), (8, 8)), (8, 8), 2)
    return (cos[64:], sin[64:]), (cos[:64], sin[:64])


def test_rng_canary(golden_dir):
    fx = torch.load(os.path.join(golden_dir, "block_tiny.pt"))["lora_rope"]
    p = O.synth_params(_cfg(fx["cfg"]), seed=fx["seed"])
    assert p.keys() == fx["params"].keys()
    for k in p:
        assert torch.equal(p[k], fx["params"][k]), k


@pytest.mark.parametrize("tag", ["lora_rope", "plain"])
def test_block_tiny(golden_dir, tag):
    fx = torch.load(os.path.join(golden_dir, "block_tiny.pt"))[tag]
    cfg = _cfg(fx["cfg"])
    p = O.synth_params(cfg, seed=fx["seed"])
    rv, rr = _rope_small() if cfg.use_rotary_positional_embeddings else (None, None)
    io = fx["io"]
    vid, txt, ref = O.block_forward(p, cfg, "transformer_blocks.0.", io["vid"], io["txt"], io["temb"], io["ref"], rv, rr)
    for got, key in ((vid, "vid"), (txt, "txt"), (ref, "ref")):
        torch.testing.assert_close(got, fx["out"][key], rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("tag", ["2b_plain", "2b_lora_rope", "5b_plain", "5b_lora_rope"])
def test_block_cfg1_shape(golden_dir, tag):
    """BASELINE.json configs[0]: single CogVideoXBlock, 1 frame 16x16 latent, random weights, CPU fp32.
    Gate from SURVEY §8d: max-abs-diff <= 1e-4."""
    fx = torch.load(os.path.join(golden_dir, "block_cfg1.pt"))[tag]
    cfg = _cfg(fx["cfg"])
    p = O.synth_params(cfg, seed=21)
    assert abs(float(sum(v.double().sum() for v in p.values())) - fx["weight_checksum"]) < 1e-6
    g = torch.Generator().manual_seed(22)
    D = cfg.inner_dim
    io = dict(vid=torch.randn(2, 64, D, generator=g), txt=torch.randn(2, 226, D, generator=g),
              ref=torch.randn(2, 64, D, generator=g), temb=torch.randn(2, 512, generator=g))
    for k in io:
        assert _sha(io[k]) == fx["io_sha"][k]
    rv, rr = _rope_small() if cfg.use_rotary_positional_embeddings else (None, None)
    outs = O.block_forward(p, cfg, "transformer_blocks.0.", io["vid"], io["txt"], io["temb"], io["ref"], rv, rr)
    for got, key in zip(outs, ("vid", "txt", "ref")):
        assert (got[..., ::16] - fx["out"][key]).abs().max() <= 1e-4
        assert abs(float(got.double().sum()) - fx["out_sum"][key]) < 1e-1


@pytest.mark.parametrize("tag", ["lora_rope", "plain_sincos"])
def test_transformer_tiny(golden_dir, tag):
    fx = torch.load(os.path.join(golden_dir, "transformer_tiny.pt"))[tag]
    cfg = _cfg(fx["cfg"])
    p = O.synth_params(cfg, seed=fx["seed"])
    io = fx["io"]
    rv = rr = None
    if cfg.use_rotary_positional_embeddings:
        rv, rr = O.pipeline_rope_tables(io["hidden"].shape[3] * 8, io["hidden"].shape[4] * 8, io["hidden"].shape[1])
    out = O.transformer_forward(p, cfg, io["hidden"], io["ref"], io["text"], io["timestep"], rv, rr, eval=True)
    torch.testing.assert_close(out, fx["out"], rtol=1e-5, atol=2e-5)


@pytest.mark.parametrize("tag,dyn", [("fp32", False), ("fp32_dyncfg", True)])
def test_pipe_loop_fp32(golden_dir, tag, dyn):
    """Reference CustomCogVideoXPipeline.__call__ (3 DDIM steps, CFG 6, 480x720, 2 latent frames) vs oracle loop."""
    fx = torch.load(os.path.join(golden_dir, "pipe_loop_tiny.pt"))
    cfg = _cfg(fx["cfg"])
    p = O.synth_params(cfg, seed=fx["seed"])
    io = fx["io"]
    pe = torch.cat([io["negative_prompt_embeds"], io["prompt_embeds"]])
    out = O.denoise_loop(p, cfg, io["latents"], pe, io["ref_img_states"], 480, 720, num_inference_steps=3,
                         guidance_scale=6.0, use_dynamic_cfg=dyn, snr_shift_scale=1.0)
    torch.testing.assert_close(out, fx["runs"][tag], rtol=1e-4, atol=1e-4)


def test_pipe_loop_bf16_reference_noise_floor(golden_dir):
    """The reference's own bf16 run differs from its fp32 run by this much — the budget any bf16 implementation has."""
    fx = torch.load(os.path.join(golden_dir, "pipe_loop_tiny.pt"))
    d = (fx["runs"]["bf16"] - fx["runs"]["fp32"]).abs()
    assert 1e-4 < float(d.max()) < 0.2


@pytest.mark.parametrize("tag,snr", [("5b", 1.0), ("2b", 3.0)])
def test_dpm_step_trace_bit_exact(golden_dir, tag, snr):
    """oracle dpm_step == the reference CogVideoXDPMScheduler CPU trace (incl. its two-draw noise protocol), bit for bit."""
    g = np.load(os.path.join(golden_dir, "scheduler_dpm.npz"))
    ac = O.ddim_alphas_cumprod(snr_shift_scale=snr)
    ts = O.ddim_timesteps(50)
    gn = torch.Generator().manual_seed(99)
    sample = torch.from_numpy(g[f"dpm_{tag}_sample0"]).to(torch.bfloat16)
    old = None
    for i, t in enumerate(ts):
        mo = torch.from_numpy(g[f"dpm_{tag}_model_out"][i])
        prev, old = O.dpm_step(ac, mo, old, int(t), int(ts[i - 1]) if i > 0 else None, sample, 50, generator=gn)
        assert np.array_equal(prev.numpy(), g[f"dpm_{tag}_prev"][i]), (tag, i)
        assert np.array_equal(old.numpy(), g[f"dpm_{tag}_x0"][i]), (tag, i)
        sample = prev.to(torch.bfloat16)


# ------------------------------------------------------------------ T5 prompt encoder oracle vs transformers' own output
def test_t5_oracle_reproduces_transformers_golden(golden_dir):
    """oracle/t5_oracle.py against tests/golden/t5_tiny.pt (transformers 5.5.0 T5EncoderModel, gated-gelu, 226 padded tokens, no
    mask): fp32 to rounding noise, and the bf16 execution bit for bit (same torch ops in the same order)."""
    from oracle import t5_oracle as T
    fx = torch.load(os.path.join(golden_dir, "t5_tiny.pt"))
    c = fx["cfg"]
    cfg = T.T5Config(d_model=c["d_model"], d_kv=c["d_kv"], d_ff=c["d_ff"], num_layers=c["num_layers"], num_heads=c["num_heads"],
                     vocab_size=c["vocab_size"])
    out = T.encoder_forward(fx["state"], cfg, fx["ids"])
    assert float((out - fx["out_fp32"]).abs().max()) <= 1e-5
    p16 = {k: v.to(torch.bfloat16) for k, v in fx["state"].items()}
    assert torch.equal(T.encoder_forward(p16, cfg, fx["ids"]).float(), fx["out_bf16"])
    # the relative-position bucket table is integer arithmetic: product copy == oracle == a few known values
    from s2v_b200 import t5
    rel = torch.arange(-300, 301)
    assert torch.equal(t5.relative_position_bucket(rel), T.relative_position_bucket(rel))
    assert T.relative_position_bucket(torch.tensor([0, 1, -1, 7, 8, -8, 127, 128, -500])).tolist() == [0, 17, 1, 23, 24, 8, 31, 31, 15]
