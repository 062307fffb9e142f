"""T5 v1.1 prompt encoder on the B200 kernels (SURVEY §8f row 3 — the step right before the denoising loop).

The reference does not implement this model: `S/inference.py:13,185-189,214` loads `transformers.T5EncoderModel` and
`D/pipelines/cogvideo/pipeline_cogvideox.py:197-237` calls `text_encoder(input_ids)[0]` on ids padded to 226 tokens, without an
attention mask.  `attach_t5(text_encoder)` binds this engine to such an ALREADY CONSTRUCTED module: its parameters are read through
transformers' state-dict names (bf16 in place; q|k|v and wi_0|wi_1 are stacked copies), `forward(input_ids)` keeps returning an object
whose `[0]` / `.last_hidden_state` is the `[B, S, d_model]` embedding, so `CustomCogVideoXPipeline._get_t5_prompt_embeds` and the
reference's `compute_prompt_embeddings` (`S/inference.py:107-140`) run unchanged.  The tokenizer stays the library's (SentencePiece).

Per layer: s2v_rmsnorm -> s2v_linear (q|k|v, no bias) -> s2v_t5_attention (relative-position bias, no 1/sqrt(d)) -> s2v_linear with the
residual add in its epilogue -> s2v_rmsnorm -> s2v_linear (wi_0|wi_1) -> s2v_gated_gelu -> s2v_linear + residual.  At 226 tokens per
prompt the 4.7 B-parameter T5-XXL is weight-bandwidth bound (9.4 GB of bf16 weights per call)."""
from __future__ import annotations

import math
import os
from typing import Dict, Optional

import torch
import torch.nn as nn

from . import _lib, ops

BF16 = torch.bfloat16


def relative_position_bucket(relative_position: torch.Tensor, num_buckets: int = 32, max_distance: int = 128) -> torch.Tensor:
    """transformers modeling_t5.py `T5Attention._relative_position_bucket`, bidirectional: integer arithmetic + one float32 log,
    evaluated with the same torch ops on the host (index arithmetic: must match bit for bit)."""
    num_buckets //= 2
    buckets = (relative_position > 0).to(torch.long) * num_buckets
    relative_position = torch.abs(relative_position)
    max_exact = num_buckets // 2
    is_small = relative_position < max_exact
    large = max_exact + (torch.log(relative_position.float() / max_exact) / math.log(max_distance / max_exact)
                         * (num_buckets - max_exact)).to(torch.long)
    large = torch.min(large, torch.full_like(large, num_buckets - 1))
    return buckets + torch.where(is_small, relative_position, large)


def _p(t: torch.Tensor) -> torch.Tensor:
    if t.dtype == torch.float16 or t.dtype == torch.float32:
        # transformers keeps `wo` in fp32 when it loads T5 in half precision; the reference's `.to(device, dtype=weight_dtype)`
        # (S/inference.py:214) casts it like everything else.  Anything not bf16 is rounded to a bf16 copy here.
        return t.detach().to(BF16).contiguous()
    if t.dtype != BF16:
        raise RuntimeError(f"T5 engine: unsupported parameter dtype {t.dtype}")
    return t.detach().contiguous()


class T5EncoderEngine:
    def __init__(self, state: Dict[str, torch.Tensor], num_layers: int, num_heads: int, d_kv: int, num_buckets: int = 32,
                 max_distance: int = 128, eps: float = 1e-6):
        if d_kv != 64:
            raise RuntimeError("s2v_t5_attention is specialised for d_kv = 64 (T5 v1.1 XXL and the other t5-v1_1 sizes)")
        self.L, self.H, self.eps = int(num_layers), int(num_heads), float(eps)
        self.num_buckets, self.max_distance = int(num_buckets), int(max_distance)
        # transformers ties encoder.embed_tokens to `shared`; the encoder's forward reads embed_tokens, so that name wins when a module
        # was built with the tie broken (e.g. materialised from the meta device)
        emb = state["encoder.embed_tokens.weight"] if "encoder.embed_tokens.weight" in state else state["shared.weight"]
        self.device = emb.device
        if self.device.type != "cuda":
            raise RuntimeError("T5EncoderEngine needs the text encoder on a CUDA (B200) device; there is no CPU path")
        self.embed = _p(emb)
        self.D = self.embed.shape[1]
        self.rel_table = _p(state["encoder.block.0.layer.0.SelfAttention.relative_attention_bias.weight"])     # [buckets, H]
        self.layers = []
        for i in range(self.L):
            a, f = f"encoder.block.{i}.layer.0.", f"encoder.block.{i}.layer.1."
            self.layers.append(dict(
                ln1=_p(state[a + "layer_norm.weight"]),
                wqkv=torch.cat([_p(state[a + f"SelfAttention.{n}.weight"]) for n in "qkv"], 0).contiguous(),
                wo=_p(state[a + "SelfAttention.o.weight"]),
                ln2=_p(state[f + "layer_norm.weight"]),
                wi=torch.cat([_p(state[f + "DenseReluDense.wi_0.weight"]), _p(state[f + "DenseReluDense.wi_1.weight"])], 0).contiguous(),
                wo_ff=_p(state[f + "DenseReluDense.wo.weight"])))
        self.final_ln = _p(state["encoder.final_layer_norm.weight"])
        self.F = self.layers[0]["wo_ff"].shape[1]
        self.inner = self.H * 64
        self.ones = torch.ones(1, self.D, device=self.device, dtype=torch.float32)     # gate table of the residual epilogue: h += 1 * y
        self._bias: Dict[int, torch.Tensor] = {}
        self._graphs: Dict[tuple, object] = {}       # (B, S) -> False (seen once) | (CUDAGraph, static ids, static output)

    def position_bias(self, S: int) -> torch.Tensor:
        """[H, S, S] bf16: relative_attention_bias gathered by bucket (T5Attention.compute_bias; no mask term — the pipeline passes none)."""
        if S not in self._bias:
            ctx = torch.arange(S, dtype=torch.long)[:, None]
            mem = torch.arange(S, dtype=torch.long)[None, :]
            bucket = relative_position_bucket(mem - ctx, self.num_buckets, self.max_distance).to(self.device)
            self._bias[S] = self.rel_table[bucket.view(-1)].view(S, S, self.H).permute(2, 0, 1).contiguous()
        return self._bias[S]

    def _call(self, name, *args):
        _lib.check(getattr(_lib.load(), name)(*args), name)

    @torch.no_grad()
    def forward(self, input_ids: torch.Tensor) -> torch.Tensor:
        """One encode.  The ~200 launches of a call are short (a 226-token prompt keeps every GEMM at the weight-streaming bound), so the
        call is launch-bound from Python: per (batch, sequence) shape the launch sequence is captured ONCE into a CUDA graph on static
        buffers and replayed (S2V_T5_GRAPH=0 keeps direct launches; the first call of a shape runs eagerly and is the warm-up)."""
        if input_ids.dim() != 2:
            raise ValueError("input_ids must be [batch, sequence]")
        ids = input_ids.to(device=self.device, dtype=torch.long).contiguous()
        if os.environ.get("S2V_T5_GRAPH", "1") == "0" or torch.cuda.is_current_stream_capturing():
            return self._forward_eager(ids)
        key = tuple(ids.shape)
        ent = self._graphs.get(key)
        if ent is None:                              # first call of this shape: eager (also loads / warms every kernel)
            self._graphs[key] = False
            return self._forward_eager(ids)
        if ent is False:                             # second call: capture
            static_ids = ids.clone()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                static_out = self._forward_eager(static_ids)
            ent = self._graphs[key] = (graph, static_ids, static_out)
        graph, static_ids, static_out = ent
        static_ids.copy_(ids)
        graph.replay()
        return static_out.clone()

    def _forward_eager(self, ids: torch.Tensor) -> torch.Tensor:
        B, S = ids.shape
        dev, D, H, M = self.device, self.D, self.H, B * S
        st = torch.cuda.current_stream().cuda_stream
        e = lambda *shape: torch.empty(*shape, device=dev, dtype=BF16)  # noqa: E731
        h, xn, qkv, att = e(M, D), e(M, D), e(M, 3 * self.inner), e(M, self.inner)
        g, act = e(M, 2 * self.F), e(M, self.F)
        bias = self.position_bias(S)
        self._call("s2v_gather_rows", self.embed.data_ptr(), ids.data_ptr(), h.data_ptr(), M, D, self.embed.shape[0], st)
        res = dict(epilogue=ops.EPI_GATE_RESIDUAL, mod=self.ones, gate_off_text=0, gate_off_other=0, rows_per_batch=M, text_len=0)
        for ly in self.layers:
            self._call("s2v_rmsnorm", h.data_ptr(), ly["ln1"].data_ptr(), xn.data_ptr(), M, D, self.eps, st)
            ops.linear(xn, ly["wqkv"], None, qkv)
            self._call("s2v_t5_attention", qkv.data_ptr(), bias.data_ptr(), att.data_ptr(), B, S, H, st)
            ops.linear(att, ly["wo"], None, h, **res)
            self._call("s2v_rmsnorm", h.data_ptr(), ly["ln2"].data_ptr(), xn.data_ptr(), M, D, self.eps, st)
            ops.linear(xn, ly["wi"], None, g)
            self._call("s2v_gated_gelu", g.data_ptr(), act.data_ptr(), M, self.F, st)
            ops.linear(act, ly["wo_ff"], None, h, **res)
        out = e(M, D)
        self._call("s2v_rmsnorm", h.data_ptr(), self.final_ln.data_ptr(), out.data_ptr(), M, D, self.eps, st)
        return out.view(B, S, D)


class T5EncoderOutput(tuple):
    """`text_encoder(ids)[0]` and `.last_hidden_state`, like transformers' BaseModelOutput."""

    def __new__(cls, last_hidden_state):
        self = super().__new__(cls, (last_hidden_state,))
        self.last_hidden_state = last_hidden_state
        return self


def attach_t5(text_encoder: nn.Module) -> nn.Module:
    """Bind the B200 engine to an already constructed `transformers.T5EncoderModel` (on a CUDA device): `text_encoder(input_ids)`
    keeps its call form and returns `T5EncoderOutput`; `text_encoder.s2v_repack()` re-reads the weights after they change.  An
    attention mask other than all-ones is refused (the reference never passes one; padding tokens attend and are attended to)."""
    cfg = text_encoder.config
    if getattr(cfg, "feed_forward_proj", "gated-gelu") != "gated-gelu":
        raise RuntimeError("only the gated-gelu (T5 v1.1) feed-forward is implemented")

    def build():
        sd = text_encoder.state_dict()
        return T5EncoderEngine(sd, cfg.num_layers, cfg.num_heads, cfg.d_kv, cfg.relative_attention_num_buckets,
                               getattr(cfg, "relative_attention_max_distance", 128), cfg.layer_norm_epsilon)

    holder = {"eng": build()}

    def forward(input_ids=None, attention_mask=None, **kw):
        if input_ids is None:
            raise ValueError("attach_t5: call with input_ids")
        if attention_mask is not None and not bool(torch.all(attention_mask == 1)):
            raise RuntimeError("attach_t5: attention masks are not implemented (the reference pipeline passes none)")
        dt = next(text_encoder.parameters()).dtype
        out = holder["eng"].forward(input_ids)
        return T5EncoderOutput(out if dt == BF16 else out.to(dt))

    def repack():
        holder["eng"] = build()

    text_encoder.forward = forward
    text_encoder.s2v_repack = repack
    text_encoder._s2v_t5 = holder
    return text_encoder
