"""Execution engine of the DiT forward on one B200: packs the (possibly LoRA-wrapped) module parameters once, owns the
activation workspaces, and drives the sm_100a kernels through the C ABI.

Data layout in HBM (per forward, B = 2 x prompts sequences):
    h    [B, S, D]   bf16  residual stream, rows [text L | ref n | video F*n]   (never split or concatenated)
    xn   [B, S, D]   bf16  AdaLN output / attention output (reused)
    qkv  [B, S, 3D]  bf16  fused projection, q|k|v along the last dim, heads contiguous -> read by TMA per head
    ffh  [B, S, 4D]  bf16  GELU(FFN up)
    lt   [B*S, 3r]   bf16  scaled LoRA down-projections  s * x A^T
    mod  [2*layers+1, B, 6D] fp32  all AdaLN-Zero modulation vectors of the step (they depend on temb only)
At cfg-3 (5B, B=2, S=19126) that is ~2.6 GB of activations next to 11 GB of weights — everything stays resident.

Per block the launch sequence is
    adaln_modulate -> [lora-down] -> qkv GEMM(+bias+LoRA) -> qk LayerNorm+RoPE -> attention
    -> [lora-down] -> out-proj GEMM(+bias+LoRA, gate*y + residual) -> adaln_modulate
    -> [lora-down] -> FFN-up GEMM(+bias+LoRA+GELU) -> [lora-down] -> FFN-down GEMM(+bias+LoRA, gate*y + residual)
replacing the ~60 library kernels and 4 full-activation cat/split copies the reference issues per block
(D/models/transformers/cogvideox_transformer_3d.py:122-186, D/models/attention_processor.py:2024-2097).
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import torch

from . import ops
from .lora import LinearParams, read_linear

BF16 = torch.bfloat16
# q/k LayerNorm + RoPE inside the QKV GEMM epilogue (s2v_qkv_lora_norm_rope); S2V_FUSED_QK_NORM=0 selects the two-kernel
# form (s2v_qkv_lora + s2v_qk_norm_rope) for A/B runs — same kernels' arithmetic, identical bits
FUSED_QK_NORM = os.environ.get("S2V_FUSED_QK_NORM", "1") != "0"


@dataclass
class PackedLinear:
    w: torch.Tensor
    b: Optional[torch.Tensor]
    a: Optional[torch.Tensor] = None   # [groups*r, K]  LoRA down (stacked per group)
    bb: Optional[torch.Tensor] = None  # [N, r]         LoRA up
    scale: float = 0.0
    group_n: int = 0

    @property
    def r(self) -> int:
        return 0 if self.bb is None else self.bb.shape[1]


_FP16_WARNED = False


def _bf16c(t: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    """Parameter as the kernels read it: bf16, contiguous.  bf16 parameters are live views.  fp16 parameters — the reference
    loads CogVideoX-2B with torch_dtype=float16 (S/inference.py:191,210) — are rounded to bf16 ONCE here (a packed copy, refreshed
    by repack()): the engine's arithmetic is bf16 x bf16 -> fp32 throughout, so a 2B run differs from the reference's fp16 run by
    bf16 rounding (8 instead of 11 significand bits; INTEGRATION.md §5 gives the measured deviation).  fp32 parameters are refused:
    nobody asking for fp32 wants a silent drop to 8 bits."""
    global _FP16_WARNED
    if t is None:
        return None
    if t.dtype == torch.float16:
        if not _FP16_WARNED:
            import warnings
            warnings.warn("s2v_b200: float16 parameters are converted to bfloat16 copies (the sm_100a kernels compute in bf16 with fp32 "
                          "accumulation); outputs are returned in float16", stacklevel=3)
            _FP16_WARNED = True
        return t.detach().to(BF16).contiguous()
    if t.dtype != BF16:
        raise RuntimeError(f"the B200 engine computes in bfloat16; got a {t.dtype} parameter (load the model with torch_dtype=bfloat16 "
                           "or float16)")
    return t.detach().contiguous()


def pack_linear(parts: List[LinearParams], merge: bool) -> PackedLinear:
    """Stack one or more (LoRA-)linears along N (q|k|v share one GEMM).  `merge` folds s*B*A into W (fp32 math, one
    rounding) instead of keeping the factors."""
    has_lora = any(p.lora_a is not None for p in parts)
    if merge and has_lora:
        ws = []
        for p in parts:
            w = p.weight.detach().float()
            if p.lora_a is not None:
                w = w + p.scale * (p.lora_b.detach().float() @ p.lora_a.detach().float())
            ws.append(w.to(BF16))
        w = torch.cat(ws, 0).contiguous() if len(ws) > 1 else ws[0].contiguous()
        has_lora = False
    else:
        w = torch.cat([_bf16c(p.weight) for p in parts], 0).contiguous() if len(parts) > 1 else _bf16c(parts[0].weight)
    if all(p.bias is None for p in parts):
        b = None
    else:
        b = torch.cat([_bf16c(p.bias) if p.bias is not None else torch.zeros(p.weight.shape[0], device=w.device, dtype=BF16)
                       for p in parts]).contiguous()
    if not has_lora:
        return PackedLinear(w, b)
    r = max(p.lora_a.shape[0] for p in parts if p.lora_a is not None)
    scale = next(p.scale for p in parts if p.lora_a is not None)
    K = parts[0].weight.shape[1]
    a_blocks, b_blocks = [], []
    for p in parts:
        n = p.weight.shape[0]
        if p.lora_a is None:
            a_blocks.append(torch.zeros(r, K, device=w.device, dtype=BF16))
            b_blocks.append(torch.zeros(n, r, device=w.device, dtype=BF16))
        else:
            if p.lora_a.shape[0] != r:
                raise RuntimeError("stacked LoRA layers must share one rank")
            # a per-layer scale different from the group's is folded into B (exact for the reference's single config)
            bw = _bf16c(p.lora_b)
            if p.scale != scale:
                bw = (bw.float() * (p.scale / scale)).to(BF16)
            a_blocks.append(_bf16c(p.lora_a))
            b_blocks.append(bw)
    a = torch.cat(a_blocks, 0).contiguous()
    bb = torch.cat(b_blocks, 0).contiguous()
    return PackedLinear(w, b, a, bb, scale, parts[0].weight.shape[0] if len(parts) > 1 else 0)


@dataclass
class PackedBlock:
    norm1: PackedLinear
    ln1_w: torch.Tensor
    ln1_b: torch.Tensor
    qkv: PackedLinear
    nq_w: torch.Tensor
    nq_b: torch.Tensor
    nk_w: torch.Tensor
    nk_b: torch.Tensor
    out: PackedLinear
    norm2: PackedLinear
    ln2_w: torch.Tensor
    ln2_b: torch.Tensor
    ff1: PackedLinear
    ff2: PackedLinear
    qk_eps: float = 1e-6
    ln_eps: float = 1e-5


def pack_block(block, merge_lora: bool = False) -> PackedBlock:
    """Read a CogVideoXBlock-shaped module (this package's or the stock diffusers one, optionally peft-adapted)."""
    at = block.attn1
    return PackedBlock(
        norm1=pack_linear([read_linear(block.norm1.linear)], merge_lora),
        ln1_w=_bf16c(block.norm1.norm.weight), ln1_b=_bf16c(block.norm1.norm.bias),
        qkv=pack_linear([read_linear(at.to_q), read_linear(at.to_k), read_linear(at.to_v)], merge_lora),
        nq_w=_bf16c(at.norm_q.weight), nq_b=_bf16c(at.norm_q.bias), nk_w=_bf16c(at.norm_k.weight), nk_b=_bf16c(at.norm_k.bias),
        out=pack_linear([read_linear(at.to_out[0])], merge_lora),
        norm2=pack_linear([read_linear(block.norm2.linear)], merge_lora),
        ln2_w=_bf16c(block.norm2.norm.weight), ln2_b=_bf16c(block.norm2.norm.bias),
        ff1=pack_linear([read_linear(block.ff.net[0].proj)], merge_lora),
        ff2=pack_linear([read_linear(block.ff.net[2])], merge_lora),
        qk_eps=float(at.norm_q.eps), ln_eps=float(block.norm1.norm.eps),
    )


class Workspace:
    """Activation buffers for one (B, S) geometry; ONE caller-side allocation of `s2v_workspace_bytes(...)` bytes carved in the
    order the header documents, reused by every block and every step (the library itself allocates nothing)."""

    def __init__(self, B: int, S: int, D: int, ff_dim: int, max_r3: int, n_mod: int, device):
        from . import _lib
        self.B, self.S, self.D = B, S, D
        total = _lib.load().s2v_workspace_bytes(B, S, D, ff_dim, max_r3, n_mod)
        if total < 0:
            raise RuntimeError(f"s2v_workspace_bytes rejected the geometry B={B} S={S} D={D} ff={ff_dim}")
        self.nbytes = int(total)
        self.arena = torch.empty(self.nbytes, device=device, dtype=torch.uint8)
        off = 0

        def carve(shape, dt=BF16):
            nonlocal off
            n = 1
            for d in shape:
                n *= d
            nb = n * (2 if dt == BF16 else 4)
            t = self.arena[off:off + nb].view(dt).view(*shape)
            off += (nb + 255) & ~255
            return t

        self.h = carve((B, S, D))
        self.xn = carve((B, S, D))
        self.att = carve((B, S, D))
        self.qkv = carve((B, S, 3 * D))
        self.ffh = carve((B, S, ff_dim))
        self.lt = carve((B * S, max(max_r3, 8)))
        self.mod = carve((n_mod, B, 6 * D), torch.float32)
        assert off == self.nbytes, (off, self.nbytes)


class BlockRunner:
    """The fused block: operates in place on the residual stream `ws.h`."""

    def __init__(self, heads: int, text_len: int):
        self.heads = heads
        self.text_len = text_len

    def _lora_down(self, pl: PackedLinear, x2d: torch.Tensor, ws: Workspace) -> Optional[torch.Tensor]:
        if pl.a is None:
            return None
        t = ws.lt[:, : pl.a.shape[0]]
        ops.linear(x2d, pl.a, None, t, alpha=pl.scale)          # t = s * x A^T   (bf16, like peft's lora_A output)
        return t

    def linear(self, pl: PackedLinear, x2d, out2d, ws: Workspace, entry="s2v_linear", **kw):
        t = self._lora_down(pl, x2d, ws)
        return ops.linear(x2d, pl.w, pl.b, out2d, lora_t=t, lora_b=pl.bb if t is not None else None,
                          lora_group_n=pl.group_n, entry=entry, **kw)

    def attention_core(self, pb: PackedBlock, ws: Workspace, x_in: torch.Tensor, rope: Optional[Tuple[torch.Tensor, torch.Tensor]]):
        """x_in [B,S,D] (normalised tokens) -> ws.att [B,S,D] = attention output BEFORE the out-projection."""
        B, S, D = x_in.shape
        cos, sin = rope if rope is not None else (None, None)
        if FUSED_QK_NORM:
            # K1+K2+K3 in one launch: LayerNorm(64) + RoPE of the q/k heads in the projection's epilogue (bit-identical to
            # the two-kernel form below, one read+write of 2/3 of qkv less)
            qk = ops.qk_norm_args(pb.nq_w, pb.nq_b, pb.nk_w, pb.nk_b, cos, sin, S, self.heads, self.text_len, pb.qk_eps)
            self.linear(pb.qkv, x_in.view(B * S, D), ws.qkv.view(B * S, 3 * D), ws, entry="s2v_qkv_lora", qk=qk)
        else:
            self.linear(pb.qkv, x_in.view(B * S, D), ws.qkv.view(B * S, 3 * D), ws, entry="s2v_qkv_lora")
            ops.qk_norm_rope(ws.qkv, pb.nq_w, pb.nq_b, pb.nk_w, pb.nk_b, cos, sin, self.heads, self.text_len, pb.qk_eps)
        ops.attention(ws.qkv, ws.att, self.heads)
        return ws.att

    def run(self, pb: PackedBlock, ws: Workspace, mod1: torch.Tensor, mod2: torch.Tensor, rope):
        """mod1/mod2: [B, 6D] fp32 = norm{1,2}.linear(silu(temb)), chunks (shift, scale, gate, enc_shift, enc_scale, enc_gate);
        video AND reference rows use chunks 0-2, text rows chunks 3-5 (D/models/normalization.py:467-484)."""
        B, S, D = ws.B, ws.S, ws.D
        L = self.text_len
        M = B * S
        ops.adaln_modulate(ws.h, ws.xn, pb.ln1_w, pb.ln1_b, mod1, shift_off_text=3 * D, scale_off_text=4 * D,
                           shift_off_other=0, scale_off_other=D, text_len=L, eps=pb.ln_eps)
        self.attention_core(pb, ws, ws.xn, rope)
        self.linear(pb.out, ws.att.view(M, D), ws.h.view(M, D), ws, entry="s2v_outproj_lora_gate_residual",
                    epilogue=ops.EPI_GATE_RESIDUAL, mod=mod1, gate_off_text=5 * D, gate_off_other=2 * D, rows_per_batch=S, text_len=L)
        ops.adaln_modulate(ws.h, ws.xn, pb.ln2_w, pb.ln2_b, mod2, shift_off_text=3 * D, scale_off_text=4 * D,
                           shift_off_other=0, scale_off_other=D, text_len=L, eps=pb.ln_eps)
        F4 = ws.ffh.shape[-1]
        self.linear(pb.ff1, ws.xn.view(M, D), ws.ffh.view(M, F4), ws, entry="s2v_ffn_up_gelu_lora", epilogue=ops.EPI_BIAS_GELU)
        self.linear(pb.ff2, ws.ffh.view(M, F4), ws.h.view(M, D), ws, entry="s2v_ffn_down_lora_gate_residual",
                    epilogue=ops.EPI_GATE_RESIDUAL, mod=mod2, gate_off_text=5 * D, gate_off_other=2 * D, rows_per_batch=S, text_len=L)


def modulation(pl: PackedLinear, x_in: torch.Tensor, out: torch.Tensor, scratch: torch.Tensor, act_in: int = 1):
    """out[B,N] = W f(x) + b (+ s B (A f(x))), fp32, f = SiLU when act_in = 1 — the 6D (or 2D) AdaLN vectors from the time
    embedding, and the two time-embedding linears themselves (any of them may carry an unmerged LoRA adapter)."""
    ops.small_linear(x_in, pl.w, pl.b, out, act_in=act_in)
    if pl.a is not None:
        u = scratch[:, : pl.a.shape[0]]
        ops.small_linear(x_in, pl.a, None, u, act_in=act_in)
        ops.small_linear(u, pl.bb, None, out, alpha=pl.scale, beta=1.0)
    return out


class TransformerEngine:
    """Whole-model forward (D/models/transformers/cogvideox_transformer_3d.py:450-560) on the fused token buffer."""

    def __init__(self, model, merge_lora: bool = False):
        cfg = model.config
        g = lambda k, d=None: (cfg[k] if isinstance(cfg, dict) else getattr(cfg, k, d))  # noqa: E731
        self.heads = int(g("num_attention_heads"))
        self.head_dim = int(g("attention_head_dim"))
        if self.head_dim != 64:
            raise RuntimeError("the sm_100a attention kernel is specialised for head_dim 64 (CogVideoX 2B/5B)")
        self.D = self.heads * self.head_dim
        self.patch = int(g("patch_size", 2))
        self.in_ch = int(g("in_channels", 16))
        self.out_ch = int(g("out_channels", 16) or 16)
        self.time_dim = int(g("time_embed_dim", 512))
        self.rotary = bool(g("use_rotary_positional_embeddings", False))
        self.spatial_scale = float(g("spatial_interpolation_scale", 1.875))
        self.temporal_scale = float(g("temporal_interpolation_scale", 1.0))
        self.freq_shift = float(g("freq_shift", 0) or 0)
        if not bool(g("flip_sin_to_cos", True)):
            raise RuntimeError("only flip_sin_to_cos=True (CogVideoX) is implemented")
        self._freqs = None
        self.model = model
        self.merge_lora = merge_lora
        self.device = next(model.parameters()).device
        if self.device.type != "cuda":
            raise RuntimeError("TransformerEngine needs the model on a CUDA (B200) device; there is no CPU path")
        self.repack()
        self._ws: Dict[Tuple[int, int], Workspace] = {}
        self._pos_cache: Dict[Tuple[int, int, int], torch.Tensor] = {}

    # ------------------------------------------------------------------ weights
    def repack(self):
        m = self.model
        ml = self.merge_lora
        self.blocks = [pack_block(b, ml) for b in m.transformer_blocks]
        pe = m.patch_embed
        self.text_proj = pack_linear([read_linear(pe.text_proj)], ml)
        self.patch_proj = pack_linear([read_linear(pe.proj)], ml)
        te = m.time_embedding
        self.t1 = pack_linear([read_linear(te.linear_1)], ml)
        self.t2 = pack_linear([read_linear(te.linear_2)], ml)
        self.nf_w, self.nf_b = _bf16c(m.norm_final.weight), _bf16c(m.norm_final.bias)
        self.no_lin = pack_linear([read_linear(m.norm_out.linear)], ml)
        self.no_w, self.no_b = _bf16c(m.norm_out.norm.weight), _bf16c(m.norm_out.norm.bias)
        self.proj_out = pack_linear([read_linear(m.proj_out)], ml)
        self.ln_eps = float(m.norm_final.eps)
        self.ff_dim = self.blocks[0].ff1.w.shape[0]
        rs = [pl.a.shape[0] for b in self.blocks for pl in (b.qkv, b.out, b.ff1, b.ff2, b.norm1, b.norm2) if pl.a is not None]
        rs += [pl.a.shape[0] for pl in (self.text_proj, self.patch_proj, self.t1, self.t2, self.no_lin, self.proj_out) if pl.a is not None]
        self.max_r = max(rs) if rs else 0
        self._modb = None      # cached modulation batches point at the packed weights

    def workspace(self, B: int, S: int) -> Workspace:
        key = (B, S)
        if key not in self._ws:
            self._ws.clear()  # one geometry at a time keeps HBM use predictable
            self._modb = None  # (its descriptors point into the old workspace)
            self._ws[key] = Workspace(B, S, self.D, self.ff_dim, self.max_r, 2 * len(self.blocks) + 1, self.device)
        return self._ws[key]

    def _modulation_batches(self, ws: Workspace, B: int):
        """The step's AdaLN modulation linears as two `s2v_small_linear_batch` problem lists (stage 1: W f(emb) + b and A f(emb) for every
        linear; stage 2: out += s B u), built once per workspace: weight and output pointers do not change between steps."""
        key = (id(ws), B)
        if self._modb is not None and self._modb[0] == key:
            return self._modb[1], self._modb[2]
        D = self.D
        pls = []
        for i, pb in enumerate(self.blocks):
            pls += [(pb.norm1, ws.mod[2 * i]), (pb.norm2, ws.mod[2 * i + 1])]
        pls.append((self.no_lin, ws.mod[2 * len(self.blocks)][:, : 2 * D]))
        n_lora = sum(1 for pl, _ in pls if pl.a is not None)
        u_all = torch.empty(max(n_lora, 1), B, max(self.max_r, 8), device=self.device, dtype=torch.float32)
        stage1, stage2 = ops.SmallLinearBatch(self.device), ops.SmallLinearBatch(self.device)
        j = 0
        for pl, out in pls:
            stage1.add(pl.w, pl.b, out, None, act_in=1)
            if pl.a is not None:
                u = u_all[j, :, : pl.a.shape[0]]
                j += 1
                stage1.add(pl.a, None, u, None, act_in=1)
                stage2.add(pl.bb, None, out, u, alpha=pl.scale, beta=1.0)
        self._modb = (key, stage1, stage2, u_all)
        return stage1, stage2

    # ------------------------------------------------------------------ forward
    def time_embed(self, timestep: torch.Tensor, B: int) -> torch.Tensor:
        """emb [B, time_dim] fp32 (sinusoid -> linear_1 -> SiLU -> linear_2; :484-491)."""
        t = timestep.to(device=self.device, dtype=torch.float32).reshape(-1)
        if t.numel() == 1 and B > 1:
            t = t.expand(B)
        t = t.contiguous()
        sin = torch.empty(B, self.D, device=self.device, dtype=torch.float32)
        if self._freqs is None:
            self._freqs = ops.timestep_freqs(self.D, self.device, self.freq_shift)
        ops.timestep_sinusoid(t, self._freqs, sin)
        h1 = torch.empty(B, self.time_dim, device=self.device, dtype=torch.float32)
        scratch = torch.empty(B, max(self.max_r, 8), device=self.device, dtype=torch.float32)
        modulation(self.t1, sin, h1, scratch, act_in=0)      # the reference's LoRA targets do not reach these two layers
        emb = torch.empty(B, self.t2.w.shape[0], device=self.device, dtype=torch.float32)
        modulation(self.t2, h1, emb, scratch, act_in=1)      # (S/inference.py:222), but another LoraConfig may
        return emb

    def _embed_rows(self, pl: PackedLinear, rows: torch.Tensor, ws: Workspace, dst_rows: List[torch.Tensor], per: int):
        """rows [len(dst)*per, K] -> each dst_rows[i] ([per, D] slice of ws.h)."""
        for i, dst in enumerate(dst_rows):
            x = rows[i * per:(i + 1) * per]
            t = None
            if pl.a is not None:
                t = ws.lt[:per, : pl.a.shape[0]]
                ops.linear(x, pl.a, None, t, alpha=pl.scale)
            ops.linear(x, pl.w, pl.b, dst, lora_t=t, lora_b=pl.bb if t is not None else None)

    def forward(self, hidden_states: torch.Tensor, ref_img_states: torch.Tensor, encoder_hidden_states: torch.Tensor,
                timestep: torch.Tensor, rope: Optional[Tuple[torch.Tensor, torch.Tensor]] = None, eval: bool = True) -> torch.Tensor:
        """hidden_states [B,F,C,H,W], ref_img_states [Br,1,C,H,W] (Br == B/2 when eval, else B), encoder_hidden_states
        [B,L,text_dim], rope = (cos, sin) fp32 [(F+1)*n, 64] in [ref | video] row order (None for non-rotary models).
        Returns the model output [B,F,C,H,W] bf16."""
        B, Fr, Cc, H, W = hidden_states.shape
        p = self.patch
        n = (H // p) * (W // p)
        L = encoder_hidden_states.shape[1]
        n_ref = ref_img_states.shape[1] * n
        S = L + n_ref + Fr * n
        D = self.D
        dev = self.device
        ws = self.workspace(B, S)
        hs = hidden_states.to(device=dev, dtype=BF16).contiguous()
        rf = ref_img_states.to(device=dev, dtype=BF16).contiguous()
        tx = encoder_hidden_states.to(device=dev, dtype=BF16).contiguous()
        Br = rf.shape[0]
        if eval and 2 * Br != B:
            raise RuntimeError(f"eval=True doubles the reference tokens along batch: need ref batch {B // 2}, got {Br}")
        if not eval and Br != B:
            raise RuntimeError(f"eval=False needs ref batch == batch ({B}), got {Br}")

        # ---- embeddings straight into the fused token buffer (text_proj computed once, not twice as at :494/:506)
        self._embed_rows(self.text_proj, tx.view(B * L, -1), ws, [ws.h[b, :L] for b in range(B)], L)
        K = Cc * p * p
        ref_rows = torch.empty(Br * n_ref, K, device=dev, dtype=BF16)
        ops.patchify(rf.view(-1, Cc, H, W), ref_rows, p)
        ref_dst = [ws.h[b, L:L + n_ref] for b in range(Br)]
        self._embed_rows(self.patch_proj, ref_rows, ws, ref_dst, n_ref)
        if eval:  # torch.cat([ref, ref], dim=0) (:503-504): both CFG halves see the same reference tokens
            for b in range(Br):
                ws.h[Br + b, L:L + n_ref].copy_(ws.h[b, L:L + n_ref])
        vid_rows = torch.empty(B * Fr * n, K, device=dev, dtype=BF16)
        ops.patchify(hs.view(-1, Cc, H, W), vid_rows, p)
        self._embed_rows(self.patch_proj, vid_rows, ws, [ws.h[b, L + n_ref:] for b in range(B)], Fr * n)
        if not self.rotary:
            ops.add_rows(ws.h, self._sincos(H // p, W // p, Fr), L + n_ref)

        # ---- all modulation vectors of this step: every block's norm1 / norm2 linear and norm_out.linear read the same time embedding, so
        # they are ONE launch (+ one for the LoRA up-projections) instead of three per linear
        emb = self.time_embed(timestep, B)
        mod_out = ws.mod[2 * len(self.blocks)][:, : 2 * D]
        if B <= 8:
            stage1, stage2 = self._modulation_batches(ws, B)
            stage1.run(emb, B)
            stage2.run(None, B)
        else:      # more than 4 prompts x 2 CFG halves: the per-linear entry point takes the rows in groups of 8
            scratch = torch.empty(B, max(self.max_r, 8), device=dev, dtype=torch.float32)
            for i, pb in enumerate(self.blocks):
                modulation(pb.norm1, emb, ws.mod[2 * i], scratch)
                modulation(pb.norm2, emb, ws.mod[2 * i + 1], scratch)
            modulation(self.no_lin, emb, mod_out, scratch)

        # ---- blocks
        if self.rotary:
            if rope is None:
                raise RuntimeError("rotary model: pass rope=(cos, sin)")
            cos, sin = rope
            if cos.shape[0] != S - L:
                raise RuntimeError(f"rope table must have {S - L} rows ([ref | video]); got {cos.shape[0]}")
            rope = (cos.to(device=dev, dtype=torch.float32).contiguous(), sin.to(device=dev, dtype=torch.float32).contiguous())
        else:
            rope = None
        runner = BlockRunner(self.heads, L)
        for i, pb in enumerate(self.blocks):
            runner.run(pb, ws, ws.mod[2 * i], ws.mod[2 * i + 1], rope)

        # ---- final norms (video rows only; the reference stream is dropped, :536-539), proj_out, unpatchify
        fn = ws.xn.view(-1)[: B * Fr * n * D].view(B, Fr * n, D)
        ops.final_norm(ws.h, fn, self.nf_w, self.nf_b, self.no_w, self.no_b, mod_out, shift_off=0, scale_off=D,
                       row0=L + n_ref, eps=self.ln_eps)
        tok = torch.empty(B * Fr * n, self.proj_out.w.shape[0], device=dev, dtype=BF16)
        runner.linear(self.proj_out, fn.view(B * Fr * n, D), tok, ws)
        out = torch.empty(B, Fr, self.out_ch, H, W, device=dev, dtype=BF16)
        ops.unpatchify(tok.view(B * Fr, n, -1), out.view(B * Fr, self.out_ch, H, W), p)
        return out

    def _sincos(self, gh: int, gw: int, frames: int) -> torch.Tensor:
        key = (gh, gw, frames)
        if key not in self._pos_cache:
            from .tables import sincos_table_3d
            t = sincos_table_3d(self.D, gw, gh, frames, self.spatial_scale, self.temporal_scale)
            self._pos_cache[key] = t.to(device=self.device, dtype=BF16).contiguous()
        return self._pos_cache[key]
