"""Drop-in boundary against the REAL reference classes.  Runs only where /root/reference exists (the build container):
imports the unmodified `CogVideoXTransformer3DModel`, `CogVideoXBlock`, `CogVideoXAttnProcessor2_0` and `CustomCogVideoXPipeline`
through tests/golden/_ref_import.py and checks that this package's mirror has the same call signatures (names, order, defaults),
the same state-dict schema, and that the engine's packing code reads the real modules — bare and with PEFT-layout LoRA layers
(tests/golden/_peft_like.py; peft itself is not installed) — in place."""
import inspect
import os
import sys

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import _ref_import  # noqa: E402

pytestmark = pytest.mark.skipif(not _ref_import.available(), reason="needs the reference tree at /root/reference (build container only)")


@pytest.fixture(scope="module")
def ref():
    _ref_import.install()
    from diffusers.models.attention_processor import CogVideoXAttnProcessor2_0
    from diffusers.models.transformers.cogvideox_transformer_3d import CogVideoXBlock, CogVideoXTransformer3DModel
    from diffusers.schedulers import CogVideoXDDIMScheduler, CogVideoXDPMScheduler
    import custom_cogvideox_pipe
    return dict(T=CogVideoXTransformer3DModel, B=CogVideoXBlock, A=CogVideoXAttnProcessor2_0, P=custom_cogvideox_pipe.CustomCogVideoXPipeline,
                DDIM=CogVideoXDDIMScheduler, DPM=CogVideoXDPMScheduler)


def _sig(fn):
    """(name, default) of every parameter except self; annotations are not compared (the mirror imports fewer typing names)."""
    return [(n, p.default if p.default is not inspect._empty else "<required>", p.kind)
            for n, p in inspect.signature(fn).parameters.items() if n != "self"]


def test_forward_signatures_equal_the_reference(ref):
    import s2v_b200
    from s2v_b200 import modules, pipeline, scheduler
    pairs = [(ref["T"].forward, modules.CogVideoXTransformer3DModel.forward),
             (ref["B"].forward, modules.CogVideoXBlock.forward),
             (ref["A"].__call__, modules.CogVideoXAttnProcessor2_0.__call__),
             (ref["P"].__call__, pipeline.CustomCogVideoXPipeline.__call__),
             (ref["DDIM"].step, scheduler.CogVideoXDDIMScheduler.step),
             (ref["DDIM"].set_timesteps, scheduler.CogVideoXDDIMScheduler.set_timesteps),
             (ref["DPM"].step, scheduler.CogVideoXDPMScheduler.step)]
    for want, got in pairs:
        assert _sig(got) == _sig(want), (want.__qualname__, _sig(got), _sig(want))
    # constructor argument names of the model (config schema) — the mirror may not add or rename any
    want = [n for n in inspect.signature(ref["T"].__init__).parameters if n != "self"]
    got = [n for n in inspect.signature(modules.CogVideoXTransformer3DModel.__init__).parameters if n != "self"]
    assert got == want
    for n, p in inspect.signature(ref["T"].__init__).parameters.items():
        if n != "self":
            assert inspect.signature(modules.CogVideoXTransformer3DModel.__init__).parameters[n].default == p.default, n
    assert s2v_b200.CustomCogVideoXPipeline._callback_tensor_inputs == ref["P"]._callback_tensor_inputs


TINY = dict(num_attention_heads=2, attention_head_dim=64, num_layers=2, time_embed_dim=64, text_embed_dim=64)


@pytest.mark.parametrize("rotary", [False, True])
def test_state_dict_schema_equals_the_reference(ref, rotary):
    from s2v_b200 import modules
    real = ref["T"](use_rotary_positional_embeddings=rotary, **TINY)
    mine = modules.CogVideoXTransformer3DModel(use_rotary_positional_embeddings=rotary, **TINY)
    want = {k: tuple(v.shape) for k, v in real.state_dict().items()}
    got = {k: tuple(v.shape) for k, v in mine.state_dict().items()}
    # the 2B patch-embed registers its sincos table as a non-persistent buffer in the reference; persistent entries must match exactly
    assert got == want
    mine.load_state_dict(real.state_dict(), strict=True)      # a reference checkpoint loads into the mirror unchanged


@pytest.mark.parametrize("with_lora", [False, True])
def test_engine_packing_reads_the_real_reference_modules_in_place(ref, with_lora):
    """pack_block / read_linear on the reference's own CogVideoXBlock (what attach() does for every block of the stock model that
    S/inference.py:191-225 builds): q|k|v stacked in reference order, LoRA factors and scale taken from the PEFT layout, single-part
    weights are views of the module's storage (no copy)."""
    import _peft_like
    from s2v_b200 import engine, lora
    torch.manual_seed(0)
    real = ref["T"](use_rotary_positional_embeddings=True, **TINY).to(torch.bfloat16)
    if with_lora:
        names = _peft_like.inject(real, 8, 4.0)
        assert set(n.rsplit(".", 1)[-1] for n in names) >= {"to_q", "to_k", "to_v", "proj", "text_proj"}
        for n in names:
            torch.nn.init.normal_(real.get_submodule(n).lora_B["default"].weight, std=0.02)
    blk = real.transformer_blocks[1]
    pb = engine.pack_block(blk)
    at = blk.attn1
    base = lambda m: getattr(m, "base_layer", m)  # noqa: E731
    D = 128
    assert torch.equal(pb.qkv.w, torch.cat([base(at.to_q).weight, base(at.to_k).weight, base(at.to_v).weight]))
    assert torch.equal(pb.qkv.b, torch.cat([base(at.to_q).bias, base(at.to_k).bias, base(at.to_v).bias]))
    assert pb.out.w.data_ptr() == base(at.to_out[0]).weight.data_ptr()               # read in place
    assert pb.ff1.w.data_ptr() == base(blk.ff.net[0].proj).weight.data_ptr()
    assert pb.ln1_w.data_ptr() == blk.norm1.norm.weight.data_ptr() and pb.nq_w.data_ptr() == at.norm_q.weight.data_ptr()
    assert pb.qk_eps == at.norm_q.eps and pb.ln_eps == blk.norm1.norm.eps
    if with_lora:
        assert pb.qkv.scale == 0.5 and pb.qkv.group_n == D
        assert torch.equal(pb.qkv.a[8:16], at.to_k.lora_A["default"].weight) and torch.equal(pb.qkv.bb[2 * D:], at.to_v.lora_B["default"].weight)
        assert pb.norm1.a.shape == (8, 64) and pb.ff2.bb.shape == (D, 8)
        # the Conv2d patch embedding with a Conv2d LoRA pair (A: 2x2 stride 2, B: 1x1) flattens to the GEMM form
        lp = lora.read_linear(real.patch_embed.proj)
        assert lp.weight.shape == (D, 16 * 4) and lp.lora_a.shape == (8, 64) and lp.lora_b.shape == (D, 8) and lp.scale == 0.5
    else:
        assert pb.qkv.a is None and pb.ff1.a is None
    # the attach() entry refuses a CPU model loudly instead of falling back
    from s2v_b200 import modules
    with pytest.raises(RuntimeError, match="CUDA"):
        modules.attach(real)
