"""Module tree with the reference's names, constructor arguments, state-dict keys and forward signatures — the
host-side mirror of the operator surface — whose forwards run on the sm_100a engine.

    CogVideoXTransformer3DModel.forward   D/models/transformers/cogvideox_transformer_3d.py:450-462
    CogVideoXBlock.forward                same file :122-136
    Attention.forward / processor         D/models/attention_processor.py:461-517, :2024-2036
    CogVideoXLayerNormZero, AdaLayerNorm  D/models/normalization.py:452-484, :28-82
    CogVideoXPatchEmbed, TimestepEmbedding D/models/embeddings.py:337-448, :831-876
    FeedForward / GELU                    D/models/attention.py:1185-1243, D/models/activations.py:65-90

The modules only HOLD parameters (so `load_state_dict` of a CogVideoX checkpoint and PEFT-layout LoRA injection work
unchanged); no forward here falls back to eager PyTorch math.  `attach(model)` binds the same engine to an
already-constructed stock diffusers model (the S/inference.py path).
"""
from __future__ import annotations

import inspect
import types
from typing import Any, Dict, Optional, Tuple, Union

import torch
import torch.nn as nn

from . import ops
from .engine import BF16, BlockRunner, TransformerEngine, Workspace, pack_block


class Transformer2DModelOutput:
    def __init__(self, sample):
        self.sample = sample


def _no_eager(name):
    def forward(self, *a, **k):
        raise RuntimeError(f"{name} holds parameters for the fused B200 engine; call the enclosing block / model instead")
    return forward


class _PackedCache:
    """Mixin for modules that cache packed weights (`_pb`) and workspaces (`_ws`): anything that replaces parameter storage
    (load_state_dict, .to() / .cuda() / .bfloat16() via _apply) drops the snapshots of the whole subtree."""

    def load_state_dict(self, *args, **kwargs):
        out = super().load_state_dict(*args, **kwargs)
        from .lora import invalidate_packed
        invalidate_packed(self)
        return out

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        from .lora import invalidate_packed
        invalidate_packed(self)
        return out


class CogVideoXLayerNormZero(nn.Module):
    def __init__(self, conditioning_dim: int, embedding_dim: int, elementwise_affine: bool = True, eps: float = 1e-5,
                 bias: bool = True):
        super().__init__()
        self.silu = nn.SiLU()
        self.linear = nn.Linear(conditioning_dim, 6 * embedding_dim, bias=bias)
        self.norm = nn.LayerNorm(embedding_dim, eps=eps, elementwise_affine=elementwise_affine)

    forward = _no_eager("CogVideoXLayerNormZero")


class AdaLayerNorm(nn.Module):
    def __init__(self, embedding_dim: int, output_dim: int, norm_elementwise_affine: bool = True, norm_eps: float = 1e-5,
                 chunk_dim: int = 1):
        super().__init__()
        self.chunk_dim = chunk_dim
        self.silu = nn.SiLU()
        self.linear = nn.Linear(embedding_dim, output_dim)
        self.norm = nn.LayerNorm(output_dim // 2, norm_eps, norm_elementwise_affine)

    forward = _no_eager("AdaLayerNorm")


class GELU(nn.Module):
    def __init__(self, dim_in: int, dim_out: int, approximate: str = "tanh", bias: bool = True):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out, bias=bias)
        self.approximate = approximate

    forward = _no_eager("GELU")


class FeedForward(nn.Module):
    def __init__(self, dim: int, inner_dim: Optional[int] = None, dropout: float = 0.0, final_dropout: bool = True,
                 bias: bool = True):
        super().__init__()
        inner_dim = inner_dim or 4 * dim
        layers = [GELU(dim, inner_dim, "tanh", bias), nn.Dropout(dropout), nn.Linear(inner_dim, dim, bias=bias)]
        if final_dropout:
            layers.append(nn.Dropout(dropout))
        self.net = nn.ModuleList(layers)

    forward = _no_eager("FeedForward")


class CogVideoXAttnProcessor2_0:
    """Same call protocol as the reference processor (attention_processor.py:2024-2036): joint attention over
    cat([encoder_hidden_states, hidden_states]) with q/k LayerNorm, RoPE on the video rows (`image_rotary_emb`) and on the
    reference-image rows [ref_img_seq_start, ref_img_seq_end) (`ref_image_rotary_emb`), out-projection, split.
    Runs qkv GEMM(+LoRA) -> qk-norm+RoPE -> tcgen05 attention -> out-proj GEMM(+LoRA) on the B200 kernels."""

    def __call__(self, attn: "Attention", hidden_states: torch.Tensor, encoder_hidden_states: torch.Tensor,
                 attention_mask: Optional[torch.Tensor] = None, image_rotary_emb=None, ref_img_seq_start: Optional[int] = 0,
                 ref_img_seq_end: Optional[int] = 0, position_delta=None, embed_ref_img: Optional[bool] = False,
                 ref_image_rotary_emb=None):
        if attention_mask is not None:
            raise RuntimeError("the B200 joint-attention kernel implements the reference's mask-free path only")
        if position_delta not in (None, 0):
            raise RuntimeError("position_delta != 0 is not used by the reference and not implemented")
        enc_len = encoder_hidden_states.size(1)
        x = torch.cat([encoder_hidden_states, hidden_states], dim=1).to(BF16).contiguous()
        B, S, D = x.shape
        pb = attn._packed()
        ws = attn._workspace(B, S, D)
        rope = None
        if image_rotary_emb is not None:
            text_len = ref_img_seq_start if embed_ref_img else enc_len
            vc, vs = image_rotary_emb
            if embed_ref_img:
                rc, rs = ref_image_rotary_emb
                cos, sin = torch.cat([rc.to(vc.device), vc]), torch.cat([rs.to(vs.device), vs])
                if ref_img_seq_end != enc_len:
                    raise RuntimeError("reference-image rows must end where the video rows start")
            else:
                cos, sin = vc, vs
            rope = (cos.to(device=x.device, dtype=torch.float32).contiguous(), sin.to(device=x.device, dtype=torch.float32).contiguous())
        else:
            text_len = enc_len
        runner = BlockRunner(attn.heads, text_len)
        runner.attention_core(pb, ws, x, rope)
        out = torch.empty_like(x)
        runner.linear(pb.out, ws.att.view(B * S, D), out.view(B * S, D), ws)
        return out[:, enc_len:], out[:, :enc_len]


class Attention(_PackedCache, nn.Module):
    def __init__(self, query_dim: int, dim_head: int = 64, heads: int = 8, qk_norm: Optional[str] = "layer_norm",
                 eps: float = 1e-6, bias: bool = True, out_bias: bool = True, dropout: float = 0.0, processor=None):
        super().__init__()
        self.inner_dim = dim_head * heads
        self.heads = heads
        self.scale = dim_head**-0.5
        self.is_cross_attention = False
        self.to_q = nn.Linear(query_dim, self.inner_dim, bias=bias)
        self.to_k = nn.Linear(query_dim, self.inner_dim, bias=bias)
        self.to_v = nn.Linear(query_dim, self.inner_dim, bias=bias)
        self.norm_q = nn.LayerNorm(dim_head, eps=eps) if qk_norm else None
        self.norm_k = nn.LayerNorm(dim_head, eps=eps) if qk_norm else None
        self.to_out = nn.ModuleList([nn.Linear(self.inner_dim, query_dim, bias=out_bias), nn.Dropout(dropout)])
        self.processor = processor or CogVideoXAttnProcessor2_0()
        self._pb = None
        self._ws = None

    def set_processor(self, processor):
        self.processor = processor

    def get_processor(self):
        return self.processor

    def _packed(self):
        if self._pb is None:
            # a stand-alone Attention has no norms/ff around it: pack only what the processor needs
            from .engine import PackedBlock, pack_linear
            from .lora import read_linear
            self._pb = PackedBlock(
                norm1=None, ln1_w=None, ln1_b=None,
                qkv=pack_linear([read_linear(self.to_q), read_linear(self.to_k), read_linear(self.to_v)], False),
                nq_w=self.norm_q.weight.detach(), nq_b=self.norm_q.bias.detach(), nk_w=self.norm_k.weight.detach(),
                nk_b=self.norm_k.bias.detach(), out=pack_linear([read_linear(self.to_out[0])], False), norm2=None, ln2_w=None,
                ln2_b=None, ff1=None, ff2=None, qk_eps=float(self.norm_q.eps))
        return self._pb

    def _workspace(self, B, S, D):
        if self._ws is None or (self._ws.B, self._ws.S) != (B, S):
            pb = self._packed()
            r3 = pb.qkv.a.shape[0] if pb.qkv.a is not None else 0
            self._ws = Workspace(B, S, D, 8, r3, 1, self.to_q.weight.device)
        return self._ws

    def forward(self, hidden_states: torch.Tensor, encoder_hidden_states: Optional[torch.Tensor] = None,
                attention_mask: Optional[torch.Tensor] = None, **cross_attention_kwargs):
        # same kwarg filtering as the reference (attention_processor.py:495-504)
        allowed = set(inspect.signature(self.processor.__call__).parameters.keys())
        kw = {k: v for k, v in cross_attention_kwargs.items() if k in allowed}
        return self.processor(self, hidden_states, encoder_hidden_states=encoder_hidden_states, attention_mask=attention_mask, **kw)


class CogVideoXBlock(_PackedCache, nn.Module):
    def __init__(self, dim: int, num_attention_heads: int, attention_head_dim: int, time_embed_dim: int, dropout: float = 0.0,
                 activation_fn: str = "gelu-approximate", attention_bias: bool = False, qk_norm: bool = True,
                 norm_elementwise_affine: bool = True, norm_eps: float = 1e-5, final_dropout: bool = True,
                 ff_inner_dim: Optional[int] = None, ff_bias: bool = True, attention_out_bias: bool = True):
        super().__init__()
        if activation_fn != "gelu-approximate":
            raise NotImplementedError("CogVideoX blocks use gelu-approximate")
        self.norm1 = CogVideoXLayerNormZero(time_embed_dim, dim, norm_elementwise_affine, norm_eps, bias=True)
        self.attn1 = Attention(query_dim=dim, dim_head=attention_head_dim, heads=num_attention_heads,
                               qk_norm="layer_norm" if qk_norm else None, eps=1e-6, bias=attention_bias, out_bias=attention_out_bias)
        self.norm2 = CogVideoXLayerNormZero(time_embed_dim, dim, norm_elementwise_affine, norm_eps, bias=True)
        self.ff = FeedForward(dim, inner_dim=ff_inner_dim, dropout=dropout, final_dropout=final_dropout, bias=ff_bias)
        self._pb = None
        self._ws = None

    def forward(self, hidden_states: torch.Tensor, encoder_hidden_states: torch.Tensor, temb: torch.Tensor,
                enc_hidden_states1: Optional[torch.Tensor] = None, image_rotary_emb=None, embed_ref_img: bool = False,
                ref_img_seq_start: Optional[int] = None, ref_img_seq_end: Optional[int] = None, position_delta=None,
                timestep=None, layer=None, ref_image_rotary_emb=None):
        """Stand-alone block call with the three streams as separate tensors (they are packed into one token buffer, the
        fused block runs in place, and views of that buffer are returned)."""
        from .engine import modulation
        if enc_hidden_states1 is None:
            raise RuntimeError("the subject-to-video block always carries the reference-image stream")
        L, n_ref = encoder_hidden_states.shape[1], enc_hidden_states1.shape[1]
        B, Nv, D = hidden_states.shape
        S = L + n_ref + Nv
        dev = hidden_states.device
        if self._pb is None:
            self._pb = pack_block(self)
        pb = self._pb
        if self._ws is None or (self._ws.B, self._ws.S) != (B, S):
            r3 = max([pl.a.shape[0] for pl in (pb.qkv, pb.out, pb.ff1, pb.ff2, pb.norm1, pb.norm2) if pl.a is not None] or [0])
            self._ws = Workspace(B, S, D, pb.ff1.w.shape[0], r3, 2, dev)
        ws = self._ws
        ws.h[:, :L].copy_(encoder_hidden_states)
        ws.h[:, L:L + n_ref].copy_(enc_hidden_states1)
        ws.h[:, L + n_ref:].copy_(hidden_states)
        emb = temb.to(device=dev, dtype=torch.float32).contiguous()
        scratch = torch.empty(B, max(self._ws.lt.shape[1], 8), device=dev, dtype=torch.float32)
        modulation(pb.norm1, emb, ws.mod[0], scratch)
        modulation(pb.norm2, emb, ws.mod[1], scratch)
        rope = None
        if image_rotary_emb is not None:
            vc, vs = image_rotary_emb
            rc, rs = ref_image_rotary_emb
            rope = (torch.cat([rc.to(dev), vc.to(dev)]).float().contiguous(), torch.cat([rs.to(dev), vs.to(dev)]).float().contiguous())
        BlockRunner(self.attn1.heads, L).run(pb, ws, ws.mod[0], ws.mod[1], rope)
        out = ws.h.clone()
        return out[:, L + n_ref:], out[:, :L], out[:, L:L + n_ref]


class CogVideoXPatchEmbed(nn.Module):
    def __init__(self, patch_size: int = 2, in_channels: int = 16, embed_dim: int = 1920, text_embed_dim: int = 4096,
                 bias: bool = True, **_unused):
        super().__init__()
        self.patch_size = patch_size
        self.proj = nn.Conv2d(in_channels, embed_dim, kernel_size=(patch_size, patch_size), stride=patch_size, bias=bias)
        self.text_proj = nn.Linear(text_embed_dim, embed_dim)

    forward = _no_eager("CogVideoXPatchEmbed")


class TimestepEmbedding(nn.Module):
    def __init__(self, in_channels: int, time_embed_dim: int):
        super().__init__()
        self.linear_1 = nn.Linear(in_channels, time_embed_dim, True)
        self.act = nn.SiLU()
        self.linear_2 = nn.Linear(time_embed_dim, time_embed_dim, True)

    forward = _no_eager("TimestepEmbedding")


class CogVideoXTransformer3DModel(_PackedCache, nn.Module):
    """CogVideoX DiT with the subject-to-video reference-image stream.  Same constructor defaults / config fields as
    cogvideox_transformer_3d.py:252-280 (2B defaults); `for_5b()` / `for_2b()` give the published geometries."""

    def __init__(self, num_attention_heads: int = 30, attention_head_dim: int = 64, in_channels: int = 16,
                 out_channels: Optional[int] = 16, flip_sin_to_cos: bool = True, freq_shift: int = 0, time_embed_dim: int = 512,
                 text_embed_dim: int = 4096, num_layers: int = 30, dropout: float = 0.0, attention_bias: bool = True,
                 sample_width: int = 90, sample_height: int = 60, sample_frames: int = 49, patch_size: int = 2,
                 temporal_compression_ratio: int = 4, max_text_seq_length: int = 226, activation_fn: str = "gelu-approximate",
                 timestep_activation_fn: str = "silu", norm_elementwise_affine: bool = True, norm_eps: float = 1e-5,
                 spatial_interpolation_scale: float = 1.875, temporal_interpolation_scale: float = 1.0,
                 use_rotary_positional_embeddings: bool = False, use_learned_positional_embeddings: bool = False):
        super().__init__()
        cfg = dict(locals())
        cfg.pop("self"); cfg.pop("__class__", None)
        self.config = types.SimpleNamespace(**cfg)
        inner = num_attention_heads * attention_head_dim
        if use_learned_positional_embeddings:
            raise ValueError("learned positional embeddings (CogVideoX-5b-I2V) are not part of the subject-to-video path")
        self.patch_embed = CogVideoXPatchEmbed(patch_size, in_channels, inner, text_embed_dim, True)
        self.embedding_dropout = nn.Dropout(dropout)
        self.time_embedding = TimestepEmbedding(inner, time_embed_dim)
        self.transformer_blocks = nn.ModuleList([
            CogVideoXBlock(dim=inner, num_attention_heads=num_attention_heads, attention_head_dim=attention_head_dim,
                           time_embed_dim=time_embed_dim, dropout=dropout, activation_fn=activation_fn,
                           attention_bias=attention_bias, norm_elementwise_affine=norm_elementwise_affine, norm_eps=norm_eps)
            for _ in range(num_layers)])
        self.norm_final = nn.LayerNorm(inner, norm_eps, norm_elementwise_affine)
        self.norm_out = AdaLayerNorm(embedding_dim=time_embed_dim, output_dim=2 * inner,
                                     norm_elementwise_affine=norm_elementwise_affine, norm_eps=norm_eps, chunk_dim=1)
        self.proj_out = nn.Linear(inner, patch_size * patch_size * (out_channels or in_channels))
        self.num_layers = num_layers
        self._engine: Optional[TransformerEngine] = None
        self.merge_lora = False

    @classmethod
    def for_5b(cls, **kw):
        return cls(num_attention_heads=48, num_layers=42, use_rotary_positional_embeddings=True, **kw)

    @classmethod
    def for_2b(cls, **kw):
        return cls(num_attention_heads=30, num_layers=30, use_rotary_positional_embeddings=False, **kw)

    @property
    def dtype(self):
        return next(self.parameters()).dtype

    @property
    def device(self):
        return next(self.parameters()).device

    # --- attention-processor protocol (cogvideox_transformer_3d.py:351-408)
    @property
    def attn_processors(self) -> Dict[str, Any]:
        return {f"transformer_blocks.{i}.attn1.processor": b.attn1.get_processor() for i, b in enumerate(self.transformer_blocks)}

    def set_attn_processor(self, processor):
        count = len(self.attn_processors)
        if isinstance(processor, dict) and len(processor) != count:
            raise ValueError(f"A dict of processors was passed, but the number of processors {len(processor)} does not match the"
                             f" number of attention layers: {count}. Please make sure to pass {count} processor classes.")
        for i, b in enumerate(self.transformer_blocks):
            b.attn1.set_processor(processor[f"transformer_blocks.{i}.attn1.processor"] if isinstance(processor, dict) else processor)

    def engine(self) -> TransformerEngine:
        if self._engine is None:
            self._engine = TransformerEngine(self, merge_lora=self.merge_lora)
        return self._engine

    def invalidate_engine(self):
        """Drop the packed-weight snapshots (engine, per-block / per-attention caches).  load_state_dict, .to()/.cuda()/.half(),
        inject_lora and load_lora_state_dict call this themselves; call it by hand after writing into parameters directly."""
        from .lora import invalidate_packed
        invalidate_packed(self)

    def forward(self, hidden_states: torch.Tensor, ref_img_states: torch.Tensor, encoder_hidden_states: torch.Tensor,
                timestep: Union[int, float, torch.LongTensor], timestep_cond: Optional[torch.Tensor] = None,
                image_rotary_emb: Optional[Tuple[torch.Tensor, torch.Tensor]] = None,
                ref_image_rotary_emb: Optional[Tuple[torch.Tensor, torch.Tensor]] = None,
                attention_kwargs: Optional[Dict[str, Any]] = None, return_dict: bool = True, eval: bool = False):
        return transformer_forward(self.engine(), hidden_states, ref_img_states, encoder_hidden_states, timestep, timestep_cond,
                                   image_rotary_emb, ref_image_rotary_emb, attention_kwargs, return_dict, eval)


def transformer_forward(engine: TransformerEngine, hidden_states, ref_img_states, encoder_hidden_states, timestep,
                        timestep_cond=None, image_rotary_emb=None, ref_image_rotary_emb=None, attention_kwargs=None,
                        return_dict=True, eval=False):
    if timestep_cond is not None:
        raise RuntimeError("timestep_cond is not used by CogVideoX and not implemented")
    if attention_kwargs is not None and attention_kwargs.get("scale", 1.0) != 1.0:
        raise RuntimeError("attention_kwargs['scale'] != 1.0 (runtime LoRA re-scaling) is not implemented; set the adapter scale instead")
    rope = None
    if image_rotary_emb is not None:
        if ref_image_rotary_emb is None:
            raise RuntimeError("rotary model: ref_image_rotary_emb is required (the reference subscripts it unconditionally)")
        dev = engine.device
        rope = (torch.cat([ref_image_rotary_emb[0].to(dev), image_rotary_emb[0].to(dev)]),
                torch.cat([ref_image_rotary_emb[1].to(dev), image_rotary_emb[1].to(dev)]))
    if not torch.is_tensor(timestep):
        timestep = torch.tensor([timestep], dtype=torch.float32)
    out = engine.forward(hidden_states, ref_img_states, encoder_hidden_states, timestep, rope=rope, eval=eval)
    if hidden_states.dtype == torch.float16:   # a float16 (CogVideoX-2B) caller gets its dtype back; the arithmetic was bf16
        out = out.to(torch.float16)
    if not return_dict:
        return (out,)
    return Transformer2DModelOutput(sample=out)


def attach(model: nn.Module, merge_lora: bool = False) -> nn.Module:
    """Bind the B200 engine to an ALREADY-CONSTRUCTED stock `diffusers` CogVideoXTransformer3DModel (optionally with peft
    LoRA layers injected, as S/inference.py:218-225 leaves it): parameters are read in place, state-dict keys are untouched,
    and `model.forward` keeps the reference signature."""
    eng = TransformerEngine(model, merge_lora=merge_lora)

    def forward(hidden_states, ref_img_states, encoder_hidden_states, timestep, timestep_cond=None, image_rotary_emb=None,
                ref_image_rotary_emb=None, attention_kwargs=None, return_dict=True, eval=False):
        return transformer_forward(eng, hidden_states, ref_img_states, encoder_hidden_states, timestep, timestep_cond,
                                   image_rotary_emb, ref_image_rotary_emb, attention_kwargs, return_dict, eval)

    model.forward = forward
    model._s2v_engine = eng
    # weights that change under the bound engine re-pack it: load_state_dict (post hook), .to()/.cuda() (instance-level _apply),
    # inject_lora / load_lora_state_dict (lora.invalidate_packed); `model.s2v_repack()` is the manual hook (e.g. after peft's
    # add_adapter / set_peft_model_state_dict, which write parameters without going through any of the above)
    model.s2v_repack = eng.repack
    model.register_load_state_dict_post_hook(lambda module, incompatible_keys: eng.repack())
    cls_apply = model._apply

    def _apply(fn, *a, **k):
        out = cls_apply(fn, *a, **k)
        eng.device = next(model.parameters()).device
        eng._ws.clear()
        eng._pos_cache.clear()
        eng._freqs = None
        eng.repack()
        return out

    model._apply = _apply
    return model
