"""Thin typed wrappers: torch CUDA tensors -> raw pointers -> the C ABI (include/s2v_b200.h).

PyTorch is used here only for device memory and streams.  Every function launches on torch's CURRENT stream,
allocates nothing unless an output tensor is not supplied, and raises RuntimeError on failure (no fallback path).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

from . import _lib
from ._lib import EPI_BIAS, EPI_BIAS_GELU, EPI_GATE_RESIDUAL, LinearArgs, QkNormArgs  # noqa: F401

BF16 = torch.bfloat16


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


class KernelTimer:
    """Optional per-kernel CUDA-event timing on the launching stream (used by bench.py for the live roofline figure).
    Events are recorded around every launch whose tag is in `names`; nothing synchronises until `summary()`."""

    def __init__(self, names):
        self.names = set(names)
        self.events = {n: [] for n in self.names}

    def summary(self):
        torch.cuda.synchronize()
        out = {}
        for n, evs in self.events.items():
            if evs:
                ms = [a.elapsed_time(b) for a, b in evs]
                out[n] = {"launches": len(ms), "avg_ms": sum(ms) / len(ms), "total_ms": sum(ms)}
        return out


_PROF: Optional[KernelTimer] = None


def set_kernel_timer(t: Optional[KernelTimer]):
    global _PROF
    _PROF = t


class _timed:
    def __init__(self, tag):
        self.on = _PROF is not None and tag in _PROF.names
        self.tag = tag

    def __enter__(self):
        if self.on:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record()

    def __exit__(self, *exc):
        if self.on:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            _PROF.events[self.tag].append((self.e0, e1))
        return False


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _chk_bf16(t: torch.Tensor, name: str):
    if not t.is_cuda or t.dtype != BF16:
        raise RuntimeError(f"{name}: expected a CUDA bfloat16 tensor, got {t.device} {t.dtype}")


def _chk_f32(t: torch.Tensor, name: str):
    if not t.is_cuda or t.dtype != torch.float32:
        raise RuntimeError(f"{name}: expected a CUDA float32 tensor, got {t.device} {t.dtype}")


def _rows(t: torch.Tensor):
    """(rows, cols, ld) of a tensor viewed as a row-major matrix over its last dim."""
    if t.stride(-1) != 1:
        raise RuntimeError("last dimension must be contiguous")
    cols = t.shape[-1]
    rows = t.numel() // cols
    if t.dim() == 1:
        return 1, cols, cols
    ld = t.stride(-2)
    # leading dims must collapse onto the row stride
    exp = ld
    for d in range(t.dim() - 2, -1, -1):
        if t.shape[d] != 1 and t.stride(d) != exp:
            raise RuntimeError("tensor is not collapsible to a 2-D row-major matrix")
        exp *= t.shape[d]
    return rows, cols, ld


def linear(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], out: torch.Tensor, *, epilogue: int = EPI_BIAS,
           alpha: float = 1.0, lora_t: Optional[torch.Tensor] = None, lora_b: Optional[torch.Tensor] = None,
           lora_group_n: int = 0, mod: Optional[torch.Tensor] = None, gate_off_text: int = 0, gate_off_other: int = 0,
           rows_per_batch: int = 0, text_len: int = 0, entry: str = "s2v_linear", qk: Optional[QkNormArgs] = None) -> torch.Tensor:
    """out[M,N] = epilogue(alpha * x[M,K] @ w[N,K]^T (+ lora_t @ lora_b^T) + bias).  See s2v_linear.
    With `qk` (built by qk_norm_args) the launch is s2v_qkv_lora_norm_rope: the q/k LayerNorm(64) + RoPE run in the epilogue."""
    for t, n in ((x, "x"), (w, "w"), (out, "out")):
        _chk_bf16(t, n)
    M, K, ldx = _rows(x)
    N, Kw, ldw = _rows(w)
    Mo, No, ldo = _rows(out)
    if Kw != K or Mo != M or No != N:
        raise RuntimeError(f"linear: shape mismatch x[{M},{K}] w[{N},{Kw}] out[{Mo},{No}]")
    a = LinearArgs()
    a.x, a.ldx, a.w, a.ldw = x.data_ptr(), ldx, w.data_ptr(), ldw
    if bias is not None:
        _chk_bf16(bias, "bias")
        if bias.numel() != N or not bias.is_contiguous():
            raise RuntimeError("linear: bias must be a contiguous [N] tensor")
    a.bias = _ptr(bias)
    if lora_t is not None:
        _chk_bf16(lora_t, "lora_t")
        _chk_bf16(lora_b, "lora_b")
        Mt, Kt, ldt = _rows(lora_t)
        Nb, r, ldb = _rows(lora_b)
        gn = lora_group_n or N
        groups = (N + gn - 1) // gn
        if Mt != M or Nb != N or Kt != groups * r:
            raise RuntimeError(f"linear: LoRA shape mismatch t[{Mt},{Kt}] b[{Nb},{r}] groups={groups}")
        a.lora_t, a.ldt, a.lora_b, a.ldb, a.lora_r, a.lora_group_n = lora_t.data_ptr(), ldt, lora_b.data_ptr(), ldb, r, gn
    else:
        a.lora_t, a.lora_b, a.lora_r, a.lora_group_n = None, None, 0, 0
    a.out, a.ldo = out.data_ptr(), ldo
    a.M, a.N, a.K, a.epilogue, a.alpha = M, N, K, epilogue, alpha
    if mod is not None:
        _chk_f32(mod, "mod")
        a.mod, a.mod_stride = mod.data_ptr(), mod.stride(0)
    else:
        a.mod, a.mod_stride = None, 0
    a.gate_off_text, a.gate_off_other, a.rows_per_batch, a.text_len = gate_off_text, gate_off_other, rows_per_batch, text_len
    lib = _lib.load()
    with _timed(entry):
        if qk is not None:
            _lib.check(lib.s2v_qkv_lora_norm_rope(C.byref(a), C.byref(qk), _stream()), "s2v_qkv_lora_norm_rope")
        else:
            _lib.check(getattr(lib, entry)(C.byref(a), _stream()), entry)
    return out


def qk_norm_args(nq_w, nq_b, nk_w, nk_b, cos: Optional[torch.Tensor], sin: Optional[torch.Tensor], S: int, heads: int,
                 text_len: int, eps: float = 1e-6) -> QkNormArgs:
    """Argument block of s2v_qkv_lora_norm_rope (the tensors must outlive the launch; the caller keeps them)."""
    for t, n in ((nq_w, "nq_w"), (nq_b, "nq_b"), (nk_w, "nk_w"), (nk_b, "nk_b")):
        _chk_bf16(t, n)
        if t.numel() != 64 or not t.is_contiguous():
            raise RuntimeError(f"qk_norm_args: {n} must be a contiguous [64] tensor")
    if (cos is None) != (sin is None):
        raise RuntimeError("qk_norm_args: cos and sin go together")
    if cos is not None:
        _chk_f32(cos, "cos"); _chk_f32(sin, "sin")
        if tuple(cos.shape) != (S - text_len, 64) or tuple(sin.shape) != (S - text_len, 64) or not cos.is_contiguous() \
                or not sin.is_contiguous():
            raise RuntimeError(f"qk_norm_args: cos/sin must be contiguous [{S - text_len},64]")
    q = QkNormArgs()
    q.nq_w, q.nq_b, q.nk_w, q.nk_b = nq_w.data_ptr(), nq_b.data_ptr(), nk_w.data_ptr(), nk_b.data_ptr()
    q.cos, q.sin = _ptr(cos), _ptr(sin)
    q.S, q.H, q.text_len, q.eps = S, heads, text_len, eps
    return q


def attention(qkv: torch.Tensor, out: torch.Tensor, heads: int, scale: Optional[float] = None) -> torch.Tensor:
    """qkv [B,S,3*H*64] -> out [B,S,H*64] (joint bidirectional attention, head_dim 64)."""
    _chk_bf16(qkv, "qkv")
    _chk_bf16(out, "out")
    B, S, W = qkv.shape
    if W != 3 * heads * 64 or not qkv.is_contiguous() or not out.is_contiguous() or tuple(out.shape) != (B, S, heads * 64):
        raise RuntimeError("attention: expected contiguous qkv [B,S,3*H*64] and out [B,S,H*64]")
    lib = _lib.load()
    with _timed("s2v_attn_fwd"):
        _lib.check(lib.s2v_attn_fwd(qkv.data_ptr(), out.data_ptr(), B, S, heads, scale if scale is not None else 0.125, _stream()),
                   "s2v_attn_fwd")
    return out


def adaln_modulate(x, out, ln_w, ln_b, mod, *, shift_off_text, scale_off_text, shift_off_other, scale_off_other,
                   text_len: int, eps: float):
    _chk_bf16(x, "x"); _chk_bf16(out, "out"); _chk_bf16(ln_w, "ln_w"); _chk_bf16(ln_b, "ln_b"); _chk_f32(mod, "mod")
    B, S, D = x.shape
    if not (x.is_contiguous() and out.is_contiguous()) or mod.shape[0] != B:
        raise RuntimeError("adaln_modulate: expected contiguous [B,S,D] tensors and mod [B, *]")
    lib = _lib.load()
    _lib.check(lib.s2v_adaln_modulate(x.data_ptr(), out.data_ptr(), ln_w.data_ptr(), ln_b.data_ptr(), mod.data_ptr(),
                                      mod.stride(0), shift_off_text, scale_off_text, shift_off_other, scale_off_other,
                                      B, S, D, text_len, eps, _stream()), "s2v_adaln_modulate")
    return out


def final_norm(x, out, w1, b1, w2, b2, mod, *, shift_off: int, scale_off: int, row0: int, eps: float):
    _chk_bf16(x, "x"); _chk_bf16(out, "out"); _chk_f32(mod, "mod")
    B, S, D = x.shape
    if not (x.is_contiguous() and out.is_contiguous()) or tuple(out.shape) != (B, S - row0, D):
        raise RuntimeError("final_norm: expected x [B,S,D], out [B,S-row0,D] contiguous")
    lib = _lib.load()
    _lib.check(lib.s2v_final_norm(x.data_ptr(), out.data_ptr(), w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr(),
                                  mod.data_ptr(), mod.stride(0), shift_off, scale_off, B, S, row0, D, eps, _stream()),
               "s2v_final_norm")
    return out


def qk_norm_rope(qkv, nq_w, nq_b, nk_w, nk_b, cos: Optional[torch.Tensor], sin: Optional[torch.Tensor], heads: int,
                 text_len: int, eps: float = 1e-6):
    _chk_bf16(qkv, "qkv")
    B, S, W = qkv.shape
    if W != 3 * heads * 64 or not qkv.is_contiguous():
        raise RuntimeError("qk_norm_rope: expected contiguous qkv [B,S,3*H*64]")
    if cos is not None:
        _chk_f32(cos, "cos"); _chk_f32(sin, "sin")
        if tuple(cos.shape) != (S - text_len, 64) or tuple(sin.shape) != (S - text_len, 64) or not cos.is_contiguous() \
                or not sin.is_contiguous():
            raise RuntimeError(f"qk_norm_rope: cos/sin must be contiguous [{S - text_len},64]")
    lib = _lib.load()
    _lib.check(lib.s2v_qk_norm_rope(qkv.data_ptr(), nq_w.data_ptr(), nq_b.data_ptr(), nk_w.data_ptr(), nk_b.data_ptr(),
                                    _ptr(cos), _ptr(sin), B, S, heads, text_len, eps, _stream()), "s2v_qk_norm_rope")
    return qkv


def small_linear(x, w, bias, out, *, act_in: int = 0, alpha: float = 1.0, beta: float = 0.0, round_bf16: bool = False):
    """out[B,N] (fp32) = beta*out + alpha*(f(x)[B,K] @ w[N,K]^T + bias)."""
    _chk_f32(x, "x"); _chk_f32(out, "out"); _chk_bf16(w, "w")
    B, K = x.shape
    N = w.shape[0]
    if w.shape[1] != K or tuple(out.shape) != (B, N) or x.stride(1) != 1 or out.stride(1) != 1 or w.stride(1) != 1:
        raise RuntimeError("small_linear: shape mismatch")
    lib = _lib.load()
    for b0 in range(0, B, 8):   # the kernel keeps up to 8 input rows in registers; larger batches (> 4 prompts x 2 CFG halves) go in groups
        nb = min(8, B - b0)
        _lib.check(lib.s2v_small_linear(x[b0:].data_ptr(), x.stride(0), w.data_ptr(), w.stride(0), _ptr(bias), out[b0:].data_ptr(),
                                        out.stride(0), nb, N, K, act_in, alpha, beta, int(round_bf16), _stream()),
                   "s2v_small_linear")
    return out


class SmallLinearBatch:
    """A fixed list of small-linear problems (`s2v_small_linear_batch`): descriptors are built once — weight, bias and output pointers do not
    change between steps — and uploaded to the device; `run(x_default, B)` is ONE launch.  Problems added with x=None read x_default."""

    def __init__(self, device):
        self.device = device
        self.items = []
        self.keep = []          # tensors whose storage the descriptors point into
        self.table = None
        self.max_n = 0

    def add(self, w, bias, out, x=None, *, act_in: int = 0, alpha: float = 1.0, beta: float = 0.0, round_bf16: bool = False):
        _chk_bf16(w, "w"); _chk_f32(out, "out")
        N, K = w.shape
        if out.shape[-1] != N or out.stride(-1) != 1 or w.stride(1) != 1 or (K % 8) or (w.stride(0) % 8):
            raise RuntimeError("SmallLinearBatch.add: shape / stride mismatch")
        if x is not None:
            _chk_f32(x, "x")
            if x.shape[-1] != K or x.stride(-1) != 1 or (x.stride(0) % 4):
                raise RuntimeError("SmallLinearBatch.add: x shape / stride mismatch")
        d = _lib.SmallLinearDesc(w.data_ptr(), w.stride(0), _ptr(bias), out.data_ptr(), out.stride(0), x.data_ptr() if x is not None else None,
                                 x.stride(0) if x is not None else 0, N, K, act_in, int(round_bf16), alpha, beta)
        self.items.append(d)
        self.keep += [t for t in (w, bias, out, x) if t is not None]
        self.max_n = max(self.max_n, N)
        self.table = None

    def __len__(self):
        return len(self.items)

    def run(self, x_default: Optional[torch.Tensor], B: int):
        if not self.items:
            return
        if B > 8:
            raise RuntimeError("SmallLinearBatch: B <= 8 per launch")
        if self.table is None:
            import ctypes as C
            arr = (_lib.SmallLinearDesc * len(self.items))(*self.items)
            raw = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)
            self.table = raw.to(self.device)
        if x_default is not None:
            _chk_f32(x_default, "x_default")
        with _timed("s2v_small_linear_batch"):
            _lib.check(_lib.load().s2v_small_linear_batch(self.table.data_ptr(), len(self.items), self.max_n, _ptr(x_default),
                                                          x_default.stride(0) if x_default is not None else 0, B, _stream()),
                       "s2v_small_linear_batch")


def timestep_freqs(dim: int, device, freq_shift: float = 0.0) -> torch.Tensor:
    """fp32 [dim/2] frequency table, evaluated on the host with the reference's expression (embeddings.py:55-61)."""
    import math
    half = dim // 2
    exponent = -math.log(10000) * torch.arange(0, half, dtype=torch.float32)
    exponent = exponent / (half - freq_shift)
    return torch.exp(exponent).to(device)


def timestep_sinusoid(t: torch.Tensor, freqs: torch.Tensor, out: torch.Tensor, round_bf16: bool = False):
    _chk_f32(t, "t"); _chk_f32(freqs, "freqs"); _chk_f32(out, "out")
    B, D = out.shape
    if freqs.numel() != D // 2 or t.numel() != B:
        raise RuntimeError("timestep_sinusoid: shape mismatch")
    lib = _lib.load()
    _lib.check(lib.s2v_timestep_sinusoid(t.data_ptr(), freqs.data_ptr(), out.data_ptr(), B, D, int(round_bf16), _stream()),
               "s2v_timestep_sinusoid")
    return out


def patchify(latents: torch.Tensor, rows: torch.Tensor, p: int):
    """latents [NB,C,H,W] bf16 -> rows [NB*(H/p)*(W/p), C*p*p]."""
    _chk_bf16(latents, "latents"); _chk_bf16(rows, "rows")
    NB, Cc, H, W = latents.shape
    if not latents.is_contiguous() or not rows.is_contiguous() or rows.numel() != latents.numel():
        raise RuntimeError("patchify: bad buffers")
    lib = _lib.load()
    _lib.check(lib.s2v_patchify(latents.data_ptr(), rows.data_ptr(), NB, Cc, H, W, p, _stream()), "s2v_patchify")
    return rows


def unpatchify(tokens: torch.Tensor, latents: torch.Tensor, p: int):
    """tokens [NB,(H/p)*(W/p),C*p*p] bf16 -> latents [NB,C,H,W]."""
    _chk_bf16(latents, "latents"); _chk_bf16(tokens, "tokens")
    NB, Cc, H, W = latents.shape
    if not latents.is_contiguous() or not tokens.is_contiguous() or tokens.numel() != latents.numel():
        raise RuntimeError("unpatchify: bad buffers")
    lib = _lib.load()
    _lib.check(lib.s2v_unpatchify(tokens.data_ptr(), latents.data_ptr(), NB, Cc, H, W, p, _stream()), "s2v_unpatchify")
    return latents


def add_rows(dst: torch.Tensor, table: torch.Tensor, row0: int):
    _chk_bf16(dst, "dst"); _chk_bf16(table, "table")
    B, S, D = dst.shape
    R = table.shape[0]
    if not dst.is_contiguous() or not table.is_contiguous() or table.shape[1] != D:
        raise RuntimeError("add_rows: bad buffers")
    lib = _lib.load()
    _lib.check(lib.s2v_add_rows(dst.data_ptr(), table.data_ptr(), B, S, D, row0, R, _stream()), "s2v_add_rows")
    return dst


def cfg_ddim_step(noise_pred: torch.Tensor, latents: torch.Tensor, latents_out: torch.Tensor, guidance: float,
                  sqrt_alpha: float, sqrt_beta: float, a_coef: float, b_coef: float, x0_out: Optional[torch.Tensor] = None):
    """noise_pred [2P,...] bf16 (uncond first), latents [P,...] bf16 -> latents_out bf16.  Bit-exact CFG + DDIM update."""
    _chk_bf16(noise_pred, "noise_pred"); _chk_bf16(latents, "latents"); _chk_bf16(latents_out, "latents_out")
    n = latents.numel()
    if noise_pred.numel() != 2 * n or latents_out.numel() != n or not (noise_pred.is_contiguous() and latents.is_contiguous()
                                                                    and latents_out.is_contiguous()):
        raise RuntimeError("cfg_ddim_step: bad buffers")
    if x0_out is not None:
        _chk_f32(x0_out, "x0_out")
    lib = _lib.load()
    _lib.check(lib.s2v_cfg_ddim_step(noise_pred.data_ptr(), latents.data_ptr(), latents_out.data_ptr(), _ptr(x0_out), n,
                                     guidance, sqrt_alpha, sqrt_beta, a_coef, b_coef, _stream()), "s2v_cfg_ddim_step")
    return latents_out


def ddim_step(model_output: torch.Tensor, sample: torch.Tensor, prev_out: torch.Tensor, sqrt_alpha: float, sqrt_beta: float,
              a_coef: float, b_coef: float, x0_out: Optional[torch.Tensor] = None):
    """fp32 model_output, bf16 sample -> fp32 prev_out (and x0_out)."""
    _chk_f32(model_output, "model_output"); _chk_bf16(sample, "sample"); _chk_f32(prev_out, "prev_out")
    n = sample.numel()
    if model_output.numel() != n or prev_out.numel() != n or not (model_output.is_contiguous() and sample.is_contiguous()
                                                                  and prev_out.is_contiguous()):
        raise RuntimeError("ddim_step: bad buffers")
    lib = _lib.load()
    _lib.check(lib.s2v_ddim_step(model_output.data_ptr(), sample.data_ptr(), prev_out.data_ptr(), _ptr(x0_out), n, sqrt_alpha,
                                 sqrt_beta, a_coef, b_coef, _stream()), "s2v_ddim_step")
    return prev_out


def video_to_uint8(video: torch.Tensor, round_half_even: bool = False) -> torch.Tensor:
    """video [B,3,F,H,W] bf16 in [-1,1] (the decoder's output) -> uint8 frames [B,F,H,W,3] on the device, with the reference's
    rounding points (s2v_video_to_uint8): truncation = export_to_video's `(frame * 255).astype(np.uint8)`, round_half_even =
    numpy_to_pil's `(images * 255).round().astype("uint8")`."""
    _chk_bf16(video, "video")
    if video.dim() != 5 or video.shape[1] != 3 or not video.is_contiguous():
        raise RuntimeError("video_to_uint8: expected a contiguous [B,3,F,H,W] bf16 tensor")
    B, _, F, H, W = video.shape
    out = torch.empty(B, F, H, W, 3, dtype=torch.uint8, device=video.device)
    lib = _lib.load()
    _lib.check(lib.s2v_video_to_uint8(video.data_ptr(), out.data_ptr(), B, F, H, W, 1 if round_half_even else 0, _stream()),
               "s2v_video_to_uint8")
    return out
