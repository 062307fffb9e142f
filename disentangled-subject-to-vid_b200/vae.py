"""3D causal VAE decoder on the sm_100a kernels, behind the reference's `AutoencoderKLCogVideoX`
decode surface (D/models/autoencoders/autoencoder_kl_cogvideox.py:984-1473, decoder half only — the encoder runs once per
video on a single reference frame and is out of scope).

    AutoencoderKLCogVideoX.decode / _decode / tiled_decode / blend_v / blend_h / enable_tiling / enable_slicing
    module tree + state-dict keys of CogVideoXDecoder3D (conv_in, mid_block.resnets.*, up_blocks.*.resnets.*,
    up_blocks.*.upsamplers.0.conv, norm_out, conv_out)            -> `load_state_dict` of a CogVideoX VAE works unchanged
    attach_vae(vae)   binds the same engine to an already constructed stock diffusers AutoencoderKLCogVideoX

Data layout: every activation is a padded channels-last VOLUME [2 + T, H + 2, W + 2, C] bf16 (include/s2v_b200.h): the
zero ring is the spatial padding of the next 3x3 convolution and the two leading frames are its causal temporal context
(the reference's conv_cache), so each convolution is ONE implicit-GEMM launch (`s2v_conv_gemm`, 27 / 9 / 1 taps) with
no im2col and no F.pad / torch.cat copies.  GroupNorm statistics, SpatialNorm + SiLU, nearest upsampling, the final
layout change and the tile seam ramps are the vectorised kernels of csrc/vae_kernels.cu.  The temporal batching
(3 + 2·k latent frames), the per-tile conv-cache chains, tile geometry and in-place blending order are the reference's
(GroupNorm statistics depend on the temporal batch, so they cannot be merged without changing results).
There is no CPU / eager fallback: the modules only hold parameters.
"""
from __future__ import annotations

import os

import ctypes as C
import math
import types
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn as nn

from . import _lib, ops
from ._lib import ConvArgs

BF16 = torch.bfloat16


class DecoderOutput:
    def __init__(self, sample):
        self.sample = sample


# ------------------------------------------------------------------------------------------------ parameter-holding modules
def _no_eager(name):
    def forward(self, *a, **k):
        raise RuntimeError(f"{name} holds parameters for the fused B200 VAE engine; call AutoencoderKLCogVideoX.decode instead")
    return forward


class CogVideoXCausalConv3d(nn.Module):
    def __init__(self, in_channels: int, out_channels: int, kernel_size: int):
        super().__init__()
        self.conv = nn.Conv3d(in_channels, out_channels, kernel_size)

    forward = _no_eager("CogVideoXCausalConv3d")


class CogVideoXSpatialNorm3D(nn.Module):
    def __init__(self, f_channels: int, zq_channels: int, groups: int = 32):
        super().__init__()
        self.norm_layer = nn.GroupNorm(num_channels=f_channels, num_groups=groups, eps=1e-6, affine=True)
        self.conv_y = CogVideoXCausalConv3d(zq_channels, f_channels, 1)
        self.conv_b = CogVideoXCausalConv3d(zq_channels, f_channels, 1)

    forward = _no_eager("CogVideoXSpatialNorm3D")


class CogVideoXResnetBlock3D(nn.Module):
    def __init__(self, in_channels: int, out_channels: int, spatial_norm_dim: int, groups: int = 32):
        super().__init__()
        self.norm1 = CogVideoXSpatialNorm3D(in_channels, spatial_norm_dim, groups)
        self.norm2 = CogVideoXSpatialNorm3D(out_channels, spatial_norm_dim, groups)
        self.conv1 = CogVideoXCausalConv3d(in_channels, out_channels, 3)
        self.conv2 = CogVideoXCausalConv3d(out_channels, out_channels, 3)
        if in_channels != out_channels:
            self.conv_shortcut = nn.Conv3d(in_channels, out_channels, 1)

    forward = _no_eager("CogVideoXResnetBlock3D")


class CogVideoXUpsample3D(nn.Module):
    def __init__(self, channels: int, compress_time: bool):
        super().__init__()
        self.conv = nn.Conv2d(channels, channels, 3, padding=1)
        self.compress_time = compress_time

    forward = _no_eager("CogVideoXUpsample3D")


class _Block(nn.Module):
    def __init__(self, cin, cout, n, z, groups, upsample: Optional[bool]):
        super().__init__()
        self.resnets = nn.ModuleList([CogVideoXResnetBlock3D(cin if i == 0 else cout, cout, z, groups) for i in range(n)])
        self.upsamplers = None if upsample is None else nn.ModuleList([CogVideoXUpsample3D(cout, upsample)])

    forward = _no_eager("CogVideoXUpBlock3D / CogVideoXMidBlock3D")


class CogVideoXDecoder3D(nn.Module):
    """Parameter tree of D/.../autoencoder_kl_cogvideox.py:842-913."""

    def __init__(self, in_channels=16, out_channels=3, block_out_channels=(128, 256, 256, 512), layers_per_block=3,
                 norm_num_groups=32, temporal_compression_ratio=4):
        super().__init__()
        ch = list(reversed(block_out_channels))
        self.conv_in = CogVideoXCausalConv3d(in_channels, ch[0], 3)
        self.mid_block = _Block(ch[0], ch[0], 2, in_channels, norm_num_groups, None)
        t_levels = int(round(math.log2(float(temporal_compression_ratio))))
        blocks, prev = [], ch[0]
        for i, c in enumerate(ch):
            last = i == len(ch) - 1
            blocks.append(_Block(prev, c, layers_per_block + 1, in_channels, norm_num_groups, None if last else (i < t_levels)))
            prev = c
        self.up_blocks = nn.ModuleList(blocks)
        self.norm_out = CogVideoXSpatialNorm3D(ch[-1], in_channels, norm_num_groups)
        self.conv_out = CogVideoXCausalConv3d(ch[-1], out_channels, 3)

    forward = _no_eager("CogVideoXDecoder3D")


class _EncResnet(nn.Module):   # CogVideoXResnetBlock3D with spatial_norm_dim=None: plain GroupNorm (A/:241-243)
    def __init__(self, in_channels: int, out_channels: int, groups: int = 32):
        super().__init__()
        self.norm1 = nn.GroupNorm(num_channels=in_channels, num_groups=groups, eps=1e-6)
        self.norm2 = nn.GroupNorm(num_channels=out_channels, num_groups=groups, eps=1e-6)
        self.conv1 = CogVideoXCausalConv3d(in_channels, out_channels, 3)
        self.conv2 = CogVideoXCausalConv3d(out_channels, out_channels, 3)
        if in_channels != out_channels:
            self.conv_shortcut = nn.Conv3d(in_channels, out_channels, 1)

    forward = _no_eager("CogVideoXResnetBlock3D")


class CogVideoXDownsample3D(nn.Module):
    def __init__(self, channels: int, compress_time: bool):
        super().__init__()
        self.conv = nn.Conv2d(channels, channels, 3, stride=2, padding=0)
        self.compress_time = compress_time

    forward = _no_eager("CogVideoXDownsample3D")


class _EncBlock(nn.Module):
    def __init__(self, cin, cout, n, groups, downsample: Optional[bool]):
        super().__init__()
        self.resnets = nn.ModuleList([_EncResnet(cin if i == 0 else cout, cout, groups) for i in range(n)])
        self.downsamplers = None if downsample is None else nn.ModuleList([CogVideoXDownsample3D(cout, downsample)])

    forward = _no_eager("CogVideoXDownBlock3D / CogVideoXMidBlock3D")


class CogVideoXEncoder3D(nn.Module):
    """Parameter tree of D/.../autoencoder_kl_cogvideox.py:682-753."""

    def __init__(self, in_channels=3, out_channels=16, block_out_channels=(128, 256, 256, 512), layers_per_block=3,
                 norm_num_groups=32, temporal_compression_ratio=4):
        super().__init__()
        ch = list(block_out_channels)
        self.conv_in = CogVideoXCausalConv3d(in_channels, ch[0], 3)
        t_levels = int(round(math.log2(float(temporal_compression_ratio))))
        blocks, prev = [], ch[0]
        for i, c in enumerate(ch):
            last = i == len(ch) - 1
            blocks.append(_EncBlock(prev, c, layers_per_block, norm_num_groups, None if last else (i < t_levels)))
            prev = c
        self.down_blocks = nn.ModuleList(blocks)
        self.mid_block = _EncBlock(ch[-1], ch[-1], 2, norm_num_groups, None)
        self.norm_out = nn.GroupNorm(norm_num_groups, ch[-1], eps=1e-6)
        self.conv_out = CogVideoXCausalConv3d(ch[-1], 2 * out_channels, 3)

    forward = _no_eager("CogVideoXEncoder3D")


# ------------------------------------------------------------------------------------------------ engine
# GroupNorm statistics of the decoder out of the producing convolution's epilogue (default); S2V_VAE_FUSED_GN=0 reads the volume with
# s2v_vae_groupnorm_stats instead (round 1; the two agree to fp32 summation order)
FUSED_GN_STATS = os.environ.get("S2V_VAE_FUSED_GN", "1") != "0"
GN_BLOCKS = 1184   # stage-1 blocks of the GroupNorm statistics (8 per SM): 296 measured 26 % of the HBM copy bandwidth


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _call(name, *args):
    lib = _lib.load()
    _lib.check(getattr(lib, name)(*args), name)


# Measurement hook (bench.py): when set to a dict, every convolution launch adds its FLOPs — "algorithmic" = 2 x output positions
# x Cout x taps x Cin over the T x H x W real positions (what a FLOP counter reports for the reference's conv3d / conv2d calls),
# "launched" = the same over the padded T x (H+2) x (W+2) rows the implicit GEMM actually computes (the border ring is waste).
CONV_FLOPS: Optional[Dict[str, float]] = None


def out_frames(T: int, compress_time: bool) -> int:
    """Frames after one CogVideoXUpsample3D (upsampling.py:385-405)."""
    if not compress_time or T == 1:
        return T
    return 1 + 2 * (T - 1) if T % 2 == 1 else 2 * T


def upsample_frame_src(T: int, compress_time: bool) -> List[int]:
    if not compress_time or T == 1:
        return list(range(T))
    if T % 2 == 1:                                   # first frame kept single, the rest doubled
        return [0] + [1 + k // 2 for k in range(2 * (T - 1))]
    return [k // 2 for k in range(2 * T)]


def spatialnorm_frame_src(T_f: int, T_l: int) -> List[int]:
    """Latent frame that F.interpolate(mode='nearest') assigns to feature frame t (autoencoder_kl_cogvideox.py:173-181)."""
    if T_f > 1 and T_f % 2 == 1:
        rest_in, rest_out = T_l - 1, T_f - 1
        return [0] + [1 + (k * rest_in) // rest_out for k in range(rest_out)]
    return [(k * T_l) // T_f for k in range(T_f)]


class _Conv:
    def __init__(self, w2d, bias, cin, cout, taps):
        self.w, self.b, self.cin, self.cout, self.taps = w2d, bias, cin, cout, taps


class _Norm:
    def __init__(self, gamma, beta, wyb, byb, C):
        self.gamma, self.beta, self.wyb, self.byb, self.C = gamma, beta, wyb, byb, C
        self.yb_off = 0      # first column of this layer's conv_y | conv_b block in the engine's concatenated table


class VaeDecoderEngine:
    """Packs the decoder parameters (state-dict names of the reference) once and runs `CogVideoXDecoder3D.forward`
    (autoencoder_kl_cogvideox.py:921-981) on padded channels-last volumes."""

    ZK = 64      # latent channels padded to one 128-byte swizzle row for the conv_y|conv_b GEMM

    def __init__(self, state: Dict[str, torch.Tensor], block_out_channels, layers_per_block: int, norm_num_groups: int,
                 temporal_compression_ratio: float, latent_channels: int = 16, prefix: str = "decoder."):
        self.ch = list(reversed(list(block_out_channels)))
        self.layers = int(layers_per_block)
        self.G = int(norm_num_groups)
        self.zc = int(latent_channels)
        self.t_levels = int(round(math.log2(float(temporal_compression_ratio))))
        p = {k[len(prefix):]: v for k, v in state.items() if k.startswith(prefix)}
        if not p:
            raise RuntimeError("no decoder.* parameters found")
        dev = next(iter(p.values())).device
        if dev.type != "cuda":
            raise RuntimeError("VaeDecoderEngine needs the VAE on a CUDA (B200) device; there is no CPU path")
        for k, v in p.items():
            if v.dtype != BF16:
                raise RuntimeError(f"the B200 VAE engine computes in bfloat16; got {v.dtype} for {k} (use vae.to(torch.bfloat16))")
        self.device = dev
        for c in self.ch:
            if c % 64 or 256 % (c // 8):
                raise RuntimeError("block_out_channels must be multiples of 64 with C/8 dividing 256 (64, 128, 256, 512)")
        self.p = p
        self.conv_in = self._conv3("conv_in", im2col=True)
        self.conv_out = self._conv3("conv_out", pad_out=8)
        self.norm_out = self._norm("norm_out")
        self.res: Dict[str, dict] = {}
        names = [f"mid_block.resnets.{i}" for i in range(2)]
        names += [f"up_blocks.{b}.resnets.{i}" for b in range(len(self.ch)) for i in range(self.layers + 1)]
        for n in names:
            r = dict(norm1=self._norm(f"{n}.norm1"), conv1=self._conv3(f"{n}.conv1"), norm2=self._norm(f"{n}.norm2"),
                     conv2=self._conv3(f"{n}.conv2"), short=None)
            if f"{n}.conv_shortcut.weight" in p:
                w = p[f"{n}.conv_shortcut.weight"]
                if w.shape[2:] != (1, 1, 1):
                    raise RuntimeError("only the 1x1x1 conv_shortcut of CogVideoX is implemented")
                r["short"] = _Conv(w.reshape(w.shape[0], w.shape[1]).contiguous(), p[f"{n}.conv_shortcut.bias"].contiguous(),
                                   w.shape[1], w.shape[0], 1)
            self.res[n] = r
        self.ups = {}
        for b in range(len(self.ch) - 1):
            w = p[f"up_blocks.{b}.upsamplers.0.conv.weight"]                        # [C, C, 3, 3]
            self.ups[b] = _Conv(w.permute(0, 2, 3, 1).reshape(w.shape[0], -1).contiguous(),
                                p[f"up_blocks.{b}.upsamplers.0.conv.bias"].contiguous(), w.shape[1], w.shape[0], 9)
        self._bufs: Dict[Tuple, torch.Tensor] = {}
        self._partial = torch.empty(GN_BLOCKS * max(self.ch) * 2, device=dev, dtype=torch.float32)
        # conv_y | conv_b of EVERY SpatialNorm of a decoder call depend only on the latent rows: one GEMM over the concatenated
        # weights produces all their tables (37 launch-bound 11 us GEMMs per call before); a layer reads its column slice
        norms = [self.res[n][k] for n in names for k in ("norm1", "norm2")] + [self.norm_out]
        off = 0
        for nm in norms:
            nm.yb_off = off
            off += 2 * nm.C
        self.yb_cols = off
        self.yb_w = torch.cat([nm.wyb for nm in norms]).contiguous()
        self.yb_b = torch.cat([nm.byb for nm in norms]).contiguous()

    # ---- packing
    def _conv3(self, name, im2col=False, pad_out=0) -> _Conv:
        w, b = self.p[f"{name}.conv.weight"], self.p[f"{name}.conv.bias"]        # [Cout, Cin, 3, 3, 3]
        cout, cin = w.shape[0], w.shape[1]
        w2 = w.permute(0, 2, 3, 4, 1).reshape(cout, 27 * cin)                     # k = ((dt*3+dh)*3+dw)*Cin + c
        if pad_out and cout < pad_out:
            w2 = torch.cat([w2, torch.zeros(pad_out - cout, w2.shape[1], device=w.device, dtype=w.dtype)])
            b = torch.cat([b, torch.zeros(pad_out - cout, device=w.device, dtype=w.dtype)])
        c = _Conv(w2.contiguous(), b.contiguous(), cin, w2.shape[0], 27)
        c.real_cout = cout
        if im2col:                      # Cin = 16 < one K block: conv_in runs as a plain GEMM over an explicit (tiny) im2col
            c.cin, c.taps = 27 * cin, 1
        return c

    def _norm(self, name) -> _Norm:
        p = self.p
        wy = p[f"{name}.conv_y.conv.weight"].reshape(-1, self.zc)
        wb = p[f"{name}.conv_b.conv.weight"].reshape(-1, self.zc)
        Cn = wy.shape[0]
        wyb = torch.zeros(2 * Cn, self.ZK, device=wy.device, dtype=BF16)
        wyb[:Cn, : self.zc], wyb[Cn:, : self.zc] = wy, wb
        byb = torch.cat([p[f"{name}.conv_y.conv.bias"], p[f"{name}.conv_b.conv.bias"]]).contiguous()
        return _Norm(p[f"{name}.norm_layer.weight"].contiguous(), p[f"{name}.norm_layer.bias"].contiguous(), wyb, byb, Cn)

    # ---- buffers
    def _vol(self, tag: str, T: int, H: int, W: int, Cn: int) -> torch.Tensor:
        key = (tag, T, H, W, Cn)
        if key not in self._bufs:
            for k in [k for k in self._bufs if k[0] == tag]:
                del self._bufs[k]                               # one geometry per tag keeps HBM use bounded
            self._bufs[key] = torch.empty(T + 2, H + 2, W + 2, Cn, device=self.device, dtype=BF16)
        return self._bufs[key]

    # ---- kernels
    def _conv(self, cv: _Conv, x: torch.Tensor, out: torch.Tensor, T: int, H: int, W: int, res: Optional[torch.Tensor] = None,
              stats: bool = False) -> Optional[torch.Tensor]:
        """One implicit-GEMM convolution.  With `stats` the GroupNorm statistics of the OUTPUT come out of the convolution's epilogue
        (s2v_conv_gemm_stats + s2v_vae_groupnorm_finalize): returns (mean, rstd) per group [G, 2] fp32 — the norm that consumes
        `out` next does not read the volume again for them."""
        a = ConvArgs()
        a.x, a.ldx, a.w, a.ldw, a.bias = x.data_ptr(), x.shape[-1], cv.w.data_ptr(), cv.w.shape[1], cv.b.data_ptr()
        a.res, a.ldres = (res.data_ptr(), res.shape[-1]) if res is not None else (None, 0)
        a.out, a.ldo = out.data_ptr(), out.shape[-1]
        a.T, a.t_pad, a.Hp, a.Wp, a.cin, a.cout, a.taps = T, 2, H + 2, W + 2, cv.cin, cv.cout, cv.taps
        if CONV_FLOPS is not None:
            per_pos = 2.0 * cv.cout * cv.taps * cv.cin
            CONV_FLOPS["algorithmic"] = CONV_FLOPS.get("algorithmic", 0.0) + per_pos * T * H * W
            CONV_FLOPS["launched"] = CONV_FLOPS.get("launched", 0.0) + per_pos * T * (H + 2) * (W + 2)
            CONV_FLOPS["launches"] = CONV_FLOPS.get("launches", 0) + 1
        if not stats:
            _call("s2v_conv_gemm", C.byref(a), _stream())
            return None
        need = ((T * (H + 2) * (W + 2) + 127) // 128) * cv.cout * 2
        if self._partial.numel() < need:
            self._partial = torch.empty(need, device=self.device, dtype=torch.float32)
        nblk = C.c_int32(0)
        _call("s2v_conv_gemm_stats", C.byref(a), self._partial.data_ptr(), C.byref(nblk), _stream())
        st = torch.empty(self.G * 2, device=self.device, dtype=torch.float32)
        _call("s2v_vae_groupnorm_finalize", self._partial.data_ptr(), st.data_ptr(), nblk.value, T, H, W, cv.cout, self.G, 1e-6, _stream())
        return st

    def _spatialnorm_silu(self, nm: _Norm, x: torch.Tensor, out: torch.Tensor, yb_all: torch.Tensor, T: int, H: int, W: int, Tl: int,
                          hl: int, wl: int, stats: Optional[torch.Tensor] = None):
        """yb_all [Tl*hl*wl, yb_cols]: conv_y | conv_b of every norm layer at latent resolution (decode_call).  `stats` = the
        (mean, rstd) table the producing convolution's epilogue left (FUSED_GN_STATS); without it the volume is read once more."""
        if stats is None:
            stats = torch.empty(self.G * 2, device=self.device, dtype=torch.float32)
            _call("s2v_vae_groupnorm_stats", x.data_ptr(), self._partial.data_ptr(), stats.data_ptr(), T, H, W, nm.C, self.G, GN_BLOCKS, 1e-6,
                  _stream())
        src = (C.c_int32 * T)(*spatialnorm_frame_src(T, Tl))
        _call("s2v_vae_spatialnorm_silu", x.data_ptr(), out.data_ptr(), stats.data_ptr(), nm.gamma.data_ptr(), nm.beta.data_ptr(),
              yb_all.data_ptr() + 2 * nm.yb_off, yb_all.shape[1], src, T, H, W, nm.C, self.G, hl, wl, _stream())

    @staticmethod
    def _context(key: str, U: torch.Tensor, T: int, cache: Dict[str, torch.Tensor], new_cache: Dict[str, torch.Tensor]):
        """Temporal context of a causal conv input volume (frames 0, 1): the previous call's last two input frames, or the
        first frame twice (autoencoder_kl_cogvideox.py:120-131).  Memory plumbing only."""
        c = cache.get(key)
        if c is None:
            U[0].copy_(U[2])
            U[1].copy_(U[2])
            c = torch.empty_like(U[:2])
        else:
            U[:2].copy_(c)
        c.copy_(U[T:T + 2])
        new_cache[key] = c

    def _resnet(self, name: str, x: torch.Tensor, xtag: str, zrows, T, H, W, Tl, hl, wl, cache, new_cache,
                xstats: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, str, Optional[torch.Tensor]]:
        r = self.res[name]
        cin, cout = r["norm1"].C, r["conv1"].cout
        u = self._vol("u", T, H, W, cin)
        self._spatialnorm_silu(r["norm1"], x, u, zrows, T, H, W, Tl, hl, wl, stats=xstats)
        self._context(f"{name}.conv1", u, T, cache, new_cache)
        h = self._vol("h", T, H, W, cout)
        hstats = self._conv(r["conv1"], u, h, T, H, W, stats=FUSED_GN_STATS)
        u2 = self._vol("u", T, H, W, cout)
        self._spatialnorm_silu(r["norm2"], h, u2, zrows, T, H, W, Tl, hl, wl, stats=hstats)
        self._context(f"{name}.conv2", u2, T, cache, new_cache)
        res = x
        if r["short"] is not None:
            res = self._vol("s", T, H, W, cout)
            self._conv(r["short"], x, res, T, H, W)
        otag = "x1" if xtag == "x0" else "x0"
        out = self._vol(otag, T, H, W, cout)
        ostats = self._conv(r["conv2"], u2, out, T, H, W, res=res, stats=FUSED_GN_STATS)     # every resnet output feeds a norm next
        return out, otag, ostats

    def decode_call(self, z: torch.Tensor, f0: int, Tl: int, i0: int, j0: int, hl: int, wl: int, scale: float,
                    cache: Dict[str, torch.Tensor], video: torch.Tensor, v0: int) -> Tuple[Dict[str, torch.Tensor], int]:
        """One `CogVideoXDecoder3D.forward` on latent frames [f0, f0+Tl) of the tile (i0, j0, hl, wl) of ONE sample
        z [C, Tz, hz, wz]; writes the decoded frames into video[:, v0:v0+n] ([3, Tv, 8*hl, 8*wl] bf16) and returns
        (new conv cache, n)."""
        Cz, Tz, hz, wz = z.shape
        new_cache: Dict[str, torch.Tensor] = {}
        zrows = torch.empty(Tl * hl * wl, self.ZK, device=self.device, dtype=BF16)
        _call("s2v_vae_latent_rows", z.data_ptr(), zrows.data_ptr(), Cz, Tz, hz, wz, f0, Tl, i0, j0, hl, wl, self.ZK, scale, _stream())
        zrows_in = zrows
        zrows = torch.empty(zrows_in.shape[0], self.yb_cols, device=self.device, dtype=BF16)     # every norm's conv_y | conv_b table
        ops.linear(zrows_in, self.yb_w, self.yb_b, zrows)
        T, H, W = Tl, hl, wl
        col = self._vol("col", T, H, W, 27 * Cz)
        _call("s2v_vae_latent_im2col", z.data_ptr(), col[2:].data_ptr(), Cz, Tz, hz, wz, f0, Tl, i0, j0, hl, wl, scale, _stream())
        x, tag = self._vol("x0", T, H, W, self.ch[0]), "x0"
        xst = self._conv(self.conv_in, col, x, T, H, W, stats=FUSED_GN_STATS)
        for i in range(2):
            x, tag, xst = self._resnet(f"mid_block.resnets.{i}", x, tag, zrows, T, H, W, Tl, hl, wl, cache, new_cache, xst)
        for b in range(len(self.ch)):
            for i in range(self.layers + 1):
                x, tag, xst = self._resnet(f"up_blocks.{b}.resnets.{i}", x, tag, zrows, T, H, W, Tl, hl, wl, cache, new_cache, xst)
            if b != len(self.ch) - 1:
                ct = b < self.t_levels
                T2 = out_frames(T, ct)
                up = self._vol("up", T2, 2 * H, 2 * W, self.ch[b])
                src = (C.c_int32 * T2)(*upsample_frame_src(T, ct))
                _call("s2v_vae_upsample_nearest", x.data_ptr(), up.data_ptr(), src, T2, H, W, self.ch[b], _stream())
                T, H, W = T2, 2 * H, 2 * W
                tag = "x1" if tag == "x0" else "x0"
                x = self._vol(tag, T, H, W, self.ch[b])
                xst = self._conv(self.ups[b], up, x, T, H, W, stats=FUSED_GN_STATS)
        u = self._vol("u", T, H, W, self.ch[-1])
        self._spatialnorm_silu(self.norm_out, x, u, zrows, T, H, W, Tl, hl, wl, stats=xst)
        self._context("conv_out", u, T, cache, new_cache)
        o = self._vol("o", T, H, W, self.conv_out.cout)
        self._conv(self.conv_out, u, o, T, H, W)
        Cv, Tv = video.shape[0], video.shape[1]
        _call("s2v_vae_volume_to_video", o.data_ptr(), video.data_ptr(), T, H, W, self.conv_out.cout, Cv, Tv, v0, _stream())
        return new_cache, T


class VaeEncoderEngine:
    """`CogVideoXEncoder3D.forward` (autoencoder_kl_cogvideox.py:755-814) for ONE frame — the reference-image path of
    S/video_generate.py:26-38 — on the same padded channels-last volumes and kernels as the decoder: implicit-GEMM causal convs
    (a single frame is its own temporal context), GroupNorm statistics + plain GroupNorm·SiLU, and the downsamplers as a
    stride-1 3x3 conv followed by the odd-position subsample (= F.pad(0,1,0,1) + stride-2 Conv2d; compress_time is the identity
    on one frame, D/models/downsampling.py:329-335).  Multi-frame video encoding is not on the inference path and is refused."""

    CPAD = 8     # image channels padded to 8 so the explicit im2col of conv_in has K = 27 * 8 (a multiple of 8)

    def __init__(self, state: Dict[str, torch.Tensor], block_out_channels, layers_per_block: int, norm_num_groups: int,
                 latent_channels: int = 16, in_channels: int = 3, prefix: str = "encoder."):
        self.ch = list(block_out_channels)
        self.layers = int(layers_per_block)
        self.G = int(norm_num_groups)
        self.zc = int(latent_channels)
        self.cin = int(in_channels)
        p = {k[len(prefix):]: v for k, v in state.items() if k.startswith(prefix)}
        if not p:
            raise RuntimeError("no encoder.* parameters found")
        dev = next(iter(p.values())).device
        if dev.type != "cuda":
            raise RuntimeError("VaeEncoderEngine needs the VAE on a CUDA (B200) device; there is no CPU path")
        for k, v in p.items():
            if v.dtype != BF16:
                raise RuntimeError(f"the B200 VAE engine computes in bfloat16; got {v.dtype} for {k} (use vae.to(torch.bfloat16))")
        if self.cin > self.CPAD:
            raise RuntimeError("at most 8 image channels")
        for c in self.ch:
            if c % 64 or 256 % (c // 8):
                raise RuntimeError("block_out_channels must be multiples of 64 with C/8 dividing 256 (64, 128, 256, 512)")
        self.device, self.p = dev, p
        w, b = p["conv_in.conv.weight"], p["conv_in.conv.bias"]                    # [C0, cin, 3, 3, 3]
        wp = torch.zeros(w.shape[0], 3, 3, 3, self.CPAD, device=dev, dtype=BF16)
        wp[..., : self.cin] = w.permute(0, 2, 3, 4, 1)
        self.conv_in = _Conv(wp.reshape(w.shape[0], 27 * self.CPAD).contiguous(), b.contiguous(), 27 * self.CPAD, w.shape[0], 1)
        self.conv_out = self._conv3("conv_out")
        self.norm_out = (p["norm_out.weight"].contiguous(), p["norm_out.bias"].contiguous())
        self.res: Dict[str, dict] = {}
        names = [f"down_blocks.{b}.resnets.{i}" for b in range(len(self.ch)) for i in range(self.layers)]
        names += [f"mid_block.resnets.{i}" for i in range(2)]
        for n in names:
            r = dict(norm1=(p[f"{n}.norm1.weight"].contiguous(), p[f"{n}.norm1.bias"].contiguous()), conv1=self._conv3(f"{n}.conv1"),
                     norm2=(p[f"{n}.norm2.weight"].contiguous(), p[f"{n}.norm2.bias"].contiguous()), conv2=self._conv3(f"{n}.conv2"),
                     short=None)
            if f"{n}.conv_shortcut.weight" in p:
                ws = p[f"{n}.conv_shortcut.weight"]
                if ws.shape[2:] != (1, 1, 1):
                    raise RuntimeError("only the 1x1x1 conv_shortcut of CogVideoX is implemented")
                r["short"] = _Conv(ws.reshape(ws.shape[0], ws.shape[1]).contiguous(), p[f"{n}.conv_shortcut.bias"].contiguous(),
                                   ws.shape[1], ws.shape[0], 1)
            self.res[n] = r
        self.downs = {}
        for b in range(len(self.ch) - 1):
            wd = p[f"down_blocks.{b}.downsamplers.0.conv.weight"]                  # [C, C, 3, 3]
            self.downs[b] = _Conv(wd.permute(0, 2, 3, 1).reshape(wd.shape[0], -1).contiguous(),
                                  p[f"down_blocks.{b}.downsamplers.0.conv.bias"].contiguous(), wd.shape[1], wd.shape[0], 9)
        self._bufs: Dict[Tuple, torch.Tensor] = {}
        self._partial = torch.empty(GN_BLOCKS * max(self.ch) * 2, device=dev, dtype=torch.float32)

    def _conv3(self, name) -> _Conv:
        w, b = self.p[f"{name}.conv.weight"], self.p[f"{name}.conv.bias"]
        return _Conv(w.permute(0, 2, 3, 4, 1).reshape(w.shape[0], 27 * w.shape[1]).contiguous(), b.contiguous(), w.shape[1], w.shape[0], 27)

    _vol = VaeDecoderEngine._vol
    _conv = VaeDecoderEngine._conv

    def _gn_silu(self, gb, x: torch.Tensor, out: torch.Tensor, H: int, W: int, Cn: int):
        stats = torch.empty(self.G * 2, device=self.device, dtype=torch.float32)
        _call("s2v_vae_groupnorm_stats", x.data_ptr(), self._partial.data_ptr(), stats.data_ptr(), 1, H, W, Cn, self.G, GN_BLOCKS, 1e-6, _stream())
        _call("s2v_vae_groupnorm_silu", x.data_ptr(), out.data_ptr(), stats.data_ptr(), gb[0].data_ptr(), gb[1].data_ptr(), 1, H, W, Cn,
              self.G, _stream())
        out[0].copy_(out[2])     # one frame: the causal context is the frame itself, twice (A/:120-127)
        out[1].copy_(out[2])

    def _resnet(self, name: str, x: torch.Tensor, xtag: str, H: int, W: int) -> Tuple[torch.Tensor, str]:
        r = self.res[name]
        cin, cout = r["norm1"][0].numel(), r["conv1"].cout
        u = self._vol("u", 1, H, W, cin)
        self._gn_silu(r["norm1"], x, u, H, W, cin)
        h = self._vol("h", 1, H, W, cout)
        self._conv(r["conv1"], u, h, 1, H, W)
        u2 = self._vol("u", 1, H, W, cout)
        self._gn_silu(r["norm2"], h, u2, H, W, cout)
        res = x
        if r["short"] is not None:
            res = self._vol("s", 1, H, W, cout)
            self._conv(r["short"], x, res, 1, H, W)
        otag = "x1" if xtag == "x0" else "x0"
        out = self._vol(otag, 1, H, W, cout)
        self._conv(r["conv2"], u2, out, 1, H, W, res=res)
        return out, otag

    def encode_call(self, img: torch.Tensor, i0: int, j0: int, ht: int, wt: int) -> torch.Tensor:
        """One encoder forward on the tile (i0, j0, ht, wt) of ONE single-frame sample img [CPAD, 1, H, W] bf16 (channels beyond
        the real ones zero) -> moments [2 * latent_channels, 1, ht // 2^(L-1), wt // 2^(L-1)] bf16."""
        Cz, Tz, Hh, Ww = img.shape
        H, W = ht, wt
        col = self._vol("col", 1, H, W, 27 * Cz)
        _call("s2v_vae_latent_im2col", img.data_ptr(), col[2:].data_ptr(), Cz, Tz, Hh, Ww, 0, 1, i0, j0, ht, wt, 1.0, _stream())
        x, tag = self._vol("x0", 1, H, W, self.ch[0]), "x0"
        self._conv(self.conv_in, col, x, 1, H, W)
        for b in range(len(self.ch)):
            for i in range(self.layers):
                x, tag = self._resnet(f"down_blocks.{b}.resnets.{i}", x, tag, H, W)
            if b != len(self.ch) - 1:
                if H < 2 or W < 2:
                    raise RuntimeError("image tile too small for the encoder's downsamplers")
                full = self._vol("dn", 1, H, W, self.ch[b])
                self._conv(self.downs[b], x, full, 1, H, W)                      # stride-1 3x3 (per-frame conv2d taps)
                Hin, Win = H, W
                H, W = Hin // 2, Win // 2
                tag = "x1" if tag == "x0" else "x0"
                x = self._vol(tag, 1, H, W, self.ch[b])
                _call("s2v_vae_subsample2", full.data_ptr(), x.data_ptr(), 1, Hin, Win, self.ch[b], _stream())
        for i in range(2):
            x, tag = self._resnet(f"mid_block.resnets.{i}", x, tag, H, W)
        u = self._vol("u", 1, H, W, self.ch[-1])
        self._gn_silu(self.norm_out, x, u, H, W, self.ch[-1])
        o = self._vol("o", 1, H, W, self.conv_out.cout)
        self._conv(self.conv_out, u, o, 1, H, W)
        moments = torch.empty(self.conv_out.cout, 1, H, W, device=self.device, dtype=BF16)
        _call("s2v_vae_volume_to_video", o.data_ptr(), moments.data_ptr(), 1, H, W, self.conv_out.cout, self.conv_out.cout, 1, 0, _stream())
        return moments


def frame_batches(num_frames: int, batch: int) -> List[Tuple[int, int]]:
    """Temporal batching of `_decode` / `tiled_decode` (autoencoder_kl_cogvideox.py:1238-1247, 1414-1419)."""
    n = max(num_frames // batch, 1)
    rem = num_frames % batch
    return [(batch * k + (0 if k == 0 else rem), min(batch * (k + 1) + rem, num_frames)) for k in range(n)]


def total_out_frames(num_frames: int, batch: int, t_levels: int) -> int:
    tot = 0
    for s, e in frame_batches(num_frames, batch):
        T = e - s
        for _ in range(t_levels):
            T = out_frames(T, True)
        tot += T
    return tot


# ------------------------------------------------------------------------------------------------ reference surface
class _DecodeMixin:
    """decode / _decode / tiled_decode / blend_* with the reference's control flow, bound to a VaeDecoderEngine."""

    def _engine(self) -> VaeDecoderEngine:
        eng = getattr(self, "_b200_engine", None)
        if eng is None:
            cfg = self.config
            g = (lambda k, d=None: cfg[k] if k in cfg else d) if isinstance(cfg, dict) else (lambda k, d=None: getattr(cfg, k, d))
            eng = VaeDecoderEngine(self.state_dict(), g("block_out_channels"), g("layers_per_block"), g("norm_num_groups"),
                                   g("temporal_compression_ratio"), g("latent_channels", 16))
            object.__setattr__(self, "_b200_engine", eng)
        return eng

    def _decode_tile(self, z1: torch.Tensor, i0: int, j0: int, hl: int, wl: int) -> torch.Tensor:
        """One sample z1 [C, T, h, w]: the conv-cache chain over the temporal batches -> [3, Tv, 8*hl, 8*wl] bf16."""
        eng = self._engine()
        nb = self.num_latent_frames_batch_size
        Tv = total_out_frames(z1.shape[1], nb, eng.t_levels)
        up = 2 ** (len(eng.ch) - 1)
        video = torch.empty(eng.conv_out.real_cout, Tv, hl * up, wl * up, device=z1.device, dtype=BF16)
        cache: Dict[str, torch.Tensor] = {}
        v0 = 0
        for s, e in frame_batches(z1.shape[1], nb):
            cache, n = eng.decode_call(z1, s, e - s, i0, j0, hl, wl, 1.0, cache, video, v0)
            v0 += n
        return video

    def _decode(self, z: torch.Tensor, return_dict: bool = True):
        b, c, t, h, w = z.shape
        if self.use_tiling and (w > self.tile_latent_min_width or h > self.tile_latent_min_height):
            return self.tiled_decode(z, return_dict=return_dict)
        zz = z.to(BF16).contiguous()
        dec = torch.stack([self._decode_tile(zz[k], 0, 0, h, w) for k in range(b)])
        return DecoderOutput(sample=dec) if return_dict else (dec,)

    def decode(self, z: torch.Tensor, return_dict: bool = True):
        if z.dim() != 5:
            raise ValueError("expected latents [batch, channels, frames, height, width]")
        if self.use_slicing and z.shape[0] > 1:
            decoded = torch.cat([self._decode(zs).sample for zs in z.split(1)])
        else:
            decoded = self._decode(z).sample
        return DecoderOutput(sample=decoded) if return_dict else (decoded,)

    @staticmethod
    def _blend(a: torch.Tensor, b: torch.Tensor, extent: int, axis: int) -> torch.Tensor:
        """In-place ramp on b along `axis` (3 = height, 4 = width) of [B, C, T, H, W] bf16 tensors."""
        extent = min(a.shape[axis], b.shape[axis], extent)
        if extent <= 0:
            return b
        if a.dtype != BF16 or b.dtype != BF16 or not a.is_cuda or not b.is_cuda:
            raise RuntimeError("blend: expected CUDA bfloat16 tensors")
        other = 4 if axis == 3 else 3
        n_other = min(a.shape[other], b.shape[other])
        if a.shape[:3] != b.shape[:3] or a.stride(0) != a.stride(1) * a.shape[1] or a.stride(1) != a.stride(2) * a.shape[2] \
                or b.stride(0) != b.stride(1) * b.shape[1] or b.stride(1) != b.stride(2) * b.shape[2]:
            raise RuntimeError("blend: leading dims must be collapsible")
        n_outer = a.shape[0] * a.shape[1] * a.shape[2]
        _call("s2v_vae_blend", a.data_ptr(), b.data_ptr(), n_outer, extent, n_other, a.shape[axis], a.stride(2), a.stride(axis),
              a.stride(other), b.stride(2), b.stride(axis), b.stride(other), _stream())
        return b

    def blend_v(self, a: torch.Tensor, b: torch.Tensor, blend_extent: int) -> torch.Tensor:
        return self._blend(a, b, blend_extent, 3)

    def blend_h(self, a: torch.Tensor, b: torch.Tensor, blend_extent: int) -> torch.Tensor:
        return self._blend(a, b, blend_extent, 4)

    def tiled_decode(self, z: torch.Tensor, return_dict: bool = True):
        """autoencoder_kl_cogvideox.py:1374-1455: same tile origins, per-tile cache chains, in-place blend order and crops."""
        b, c, t, height, width = z.shape
        oh = int(self.tile_latent_min_height * (1 - self.tile_overlap_factor_height))
        ow = int(self.tile_latent_min_width * (1 - self.tile_overlap_factor_width))
        bh = int(self.tile_sample_min_height * self.tile_overlap_factor_height)
        bw = int(self.tile_sample_min_width * self.tile_overlap_factor_width)
        lh, lw = self.tile_sample_min_height - bh, self.tile_sample_min_width - bw
        zz = z.to(BF16).contiguous()
        rows = []
        for i in range(0, height, oh):
            row = []
            for j in range(0, width, ow):
                hl, wl = min(self.tile_latent_min_height, height - i), min(self.tile_latent_min_width, width - j)
                row.append(torch.stack([self._decode_tile(zz[k], i, j, hl, wl) for k in range(b)]))
            rows.append(row)
        result_rows = []
        for i, row in enumerate(rows):
            result_row = []
            for j, tile in enumerate(row):
                if i > 0:
                    tile = self.blend_v(rows[i - 1][j], tile, bh)
                if j > 0:
                    tile = self.blend_h(row[j - 1], tile, bw)
                result_row.append(tile[:, :, :, :lh, :lw])
            result_rows.append(torch.cat(result_row, dim=4))
        dec = torch.cat(result_rows, dim=3)
        return DecoderOutput(sample=dec) if return_dict else (dec,)


class AutoencoderKLOutput:
    def __init__(self, latent_dist):
        self.latent_dist = latent_dist


class DiagonalGaussianDistribution:
    """D/models/autoencoders/vae.py:767-820 (the members the inference path touches): mean | logvar moments, clamped logvar,
    `sample(generator)` drawing with torch's generator exactly like randn_tensor, `mode()`."""

    def __init__(self, parameters: torch.Tensor, deterministic: bool = False):
        self.parameters = parameters
        self.mean, self.logvar = torch.chunk(parameters, 2, dim=1)
        self.logvar = torch.clamp(self.logvar, -30.0, 20.0)
        self.deterministic = deterministic
        self.std = torch.exp(0.5 * self.logvar)
        self.var = torch.exp(self.logvar)
        if deterministic:
            self.var = self.std = torch.zeros_like(self.mean)

    def sample(self, generator: Optional[torch.Generator] = None) -> torch.Tensor:
        gdev = generator.device if generator is not None else self.parameters.device
        noise = torch.randn(self.mean.shape, generator=generator, device=gdev, dtype=self.parameters.dtype).to(self.parameters.device)
        return self.mean + self.std * noise

    def mode(self) -> torch.Tensor:
        return self.mean


class _EncodeMixin:
    """encode / _encode / tiled_encode with the reference's control flow (autoencoder_kl_cogvideox.py:1177-1229, 1300-1372),
    bound to a VaeEncoderEngine; single-frame inputs only (the reference-image path)."""

    def _enc_engine(self) -> VaeEncoderEngine:
        eng = getattr(self, "_b200_enc_engine", None)
        if eng is None:
            cfg = self.config
            g = (lambda k, d=None: cfg[k] if k in cfg else d) if isinstance(cfg, dict) else (lambda k, d=None: getattr(cfg, k, d))
            eng = VaeEncoderEngine(self.state_dict(), g("block_out_channels"), g("layers_per_block"), g("norm_num_groups"),
                                   g("latent_channels", 16), g("in_channels", 3))
            object.__setattr__(self, "_b200_enc_engine", eng)
        return eng

    def _encode_image(self, x1: torch.Tensor) -> torch.Tensor:
        """x1 [C, 1, H, W] -> the engine's channel-padded bf16 image."""
        eng = self._enc_engine()
        img = torch.zeros(eng.CPAD, 1, x1.shape[2], x1.shape[3], device=x1.device, dtype=BF16)
        img[: x1.shape[0]] = x1.to(BF16)
        return img

    def _encode(self, x: torch.Tensor) -> torch.Tensor:
        b, c, t, h, w = x.shape
        if t != 1:
            raise NotImplementedError("the B200 VAE encoder covers the single-frame reference-image path (S/video_generate.py:26-38); "
                                      "multi-frame video encoding is not on the inference path")
        if self.use_tiling and (w > self.tile_sample_min_width or h > self.tile_sample_min_height):
            return self.tiled_encode(x)
        eng = self._enc_engine()
        return torch.stack([eng.encode_call(self._encode_image(x[k]), 0, 0, h, w) for k in range(b)])

    def encode(self, x: torch.Tensor, return_dict: bool = True):
        if x.dim() != 5:
            raise ValueError("expected images [batch, channels, frames, height, width]")
        if self.use_slicing and x.shape[0] > 1:
            h = torch.cat([self._encode(xs) for xs in x.split(1)])
        else:
            h = self._encode(x)
        posterior = DiagonalGaussianDistribution(h)
        return AutoencoderKLOutput(latent_dist=posterior) if return_dict else (posterior,)

    def tiled_encode(self, x: torch.Tensor) -> torch.Tensor:
        b, c, t, height, width = x.shape
        if t != 1:
            raise NotImplementedError("the B200 VAE encoder covers the single-frame reference-image path")
        overlap_height = int(self.tile_sample_min_height * (1 - self.tile_overlap_factor_height))
        overlap_width = int(self.tile_sample_min_width * (1 - self.tile_overlap_factor_width))
        blend_extent_height = int(self.tile_latent_min_height * self.tile_overlap_factor_height)
        blend_extent_width = int(self.tile_latent_min_width * self.tile_overlap_factor_width)
        row_limit_height = self.tile_latent_min_height - blend_extent_height
        row_limit_width = self.tile_latent_min_width - blend_extent_width
        eng = self._enc_engine()
        imgs = [self._encode_image(x[k]) for k in range(b)]
        rows = []
        for i in range(0, height, overlap_height):
            row = []
            for j in range(0, width, overlap_width):
                ht, wt = min(self.tile_sample_min_height, height - i), min(self.tile_sample_min_width, width - j)
                row.append(torch.stack([eng.encode_call(im, i, j, ht, wt) for im in imgs]))
            rows.append(row)
        result_rows = []
        for i, row in enumerate(rows):
            result_row = []
            for j, tile in enumerate(row):
                if i > 0:
                    tile = self.blend_v(rows[i - 1][j], tile, blend_extent_height)
                if j > 0:
                    tile = self.blend_h(row[j - 1], tile, blend_extent_width)
                result_row.append(tile[:, :, :, :row_limit_height, :row_limit_width])
            result_rows.append(torch.cat(result_row, dim=4))
        return torch.cat(result_rows, dim=3)


def encode_reference_image(vae, image, generator: Optional[torch.Generator] = None, dtype=BF16) -> torch.Tensor:
    """S/video_generate.py:26-38: RGB image (PIL image or uint8 array [H, W, 3]) -> `ref_img_states` [1, 1, C, h, w]:
    float / 255 * 2 - 1, [1, 3, 1, H, W] in the transformer dtype, vae.encode(...).latent_dist.sample() * scaling_factor, frames
    moved in front of channels."""
    import numpy as np

    arr = np.array(image.convert("RGB")) if hasattr(image, "convert") else np.asarray(image)
    if arr.dtype != np.uint8 or arr.ndim != 3 or arr.shape[2] != 3:
        raise ValueError("expected an RGB uint8 image [H, W, 3]")
    dev = next(vae.parameters()).device
    x = torch.from_numpy(np.expand_dims(arr, axis=0)).float() / 255.0 * 2.0 - 1.0
    x = x.permute(0, 3, 1, 2).unsqueeze(0).permute(0, 2, 1, 3, 4).to(device=dev, dtype=dtype)
    cfg = vae.config
    scaling = cfg["scaling_factor"] if isinstance(cfg, dict) else cfg.scaling_factor
    lat = vae.encode(x).latent_dist.sample(generator) * scaling
    return lat.permute(0, 2, 1, 3, 4)


def _init_tiling(self, sample_height: int, sample_width: int, n_levels: int):
    """Tiling constants of AutoencoderKLCogVideoX.__init__ (autoencoder_kl_cogvideox.py:1095-1115)."""
    self.use_slicing = False
    self.use_tiling = False
    self.num_latent_frames_batch_size = 2
    self.num_sample_frames_batch_size = 8
    self.tile_sample_min_height = sample_height // 2
    self.tile_sample_min_width = sample_width // 2
    self.tile_latent_min_height = int(self.tile_sample_min_height / (2 ** (n_levels - 1)))
    self.tile_latent_min_width = int(self.tile_sample_min_width / (2 ** (n_levels - 1)))
    self.tile_overlap_factor_height = 1 / 6
    self.tile_overlap_factor_width = 1 / 5


class AutoencoderKLCogVideoX(_DecodeMixin, _EncodeMixin, nn.Module):
    """The reference class with the same constructor arguments, config fields, state-dict keys and decode surface; `encode`
    covers single-frame inputs (the reference image, S/video_generate.py:26-38)."""

    def __init__(self, in_channels: int = 3, out_channels: int = 3, block_out_channels=(128, 256, 256, 512), latent_channels: int = 16,
                 layers_per_block: int = 3, act_fn: str = "silu", norm_eps: float = 1e-6, norm_num_groups: int = 32,
                 temporal_compression_ratio: float = 4, sample_height: int = 480, sample_width: int = 720,
                 scaling_factor: float = 1.15258426, **unused):
        super().__init__()
        if act_fn != "silu" or norm_eps != 1e-6:
            raise NotImplementedError("the B200 VAE engine implements CogVideoX's SiLU / eps 1e-6 decoder")
        self.config = types.SimpleNamespace(
            in_channels=in_channels, out_channels=out_channels, block_out_channels=tuple(block_out_channels),
            latent_channels=latent_channels, layers_per_block=layers_per_block, act_fn=act_fn, norm_eps=norm_eps,
            norm_num_groups=norm_num_groups, temporal_compression_ratio=temporal_compression_ratio, sample_height=sample_height,
            sample_width=sample_width, scaling_factor=scaling_factor)
        self.encoder = CogVideoXEncoder3D(in_channels, latent_channels, block_out_channels, layers_per_block, norm_num_groups,
                                          temporal_compression_ratio)
        self.decoder = CogVideoXDecoder3D(latent_channels, out_channels, block_out_channels, layers_per_block, norm_num_groups,
                                          temporal_compression_ratio)
        self.quant_conv = None
        self.post_quant_conv = None
        _init_tiling(self, sample_height, sample_width, len(block_out_channels))

    def enable_tiling(self, tile_sample_min_height=None, tile_sample_min_width=None, tile_overlap_factor_height=None,
                      tile_overlap_factor_width=None) -> None:
        self.use_tiling = True
        self.tile_sample_min_height = tile_sample_min_height or self.tile_sample_min_height
        self.tile_sample_min_width = tile_sample_min_width or self.tile_sample_min_width
        self.tile_latent_min_height = int(self.tile_sample_min_height / (2 ** (len(self.config.block_out_channels) - 1)))
        self.tile_latent_min_width = int(self.tile_sample_min_width / (2 ** (len(self.config.block_out_channels) - 1)))
        self.tile_overlap_factor_height = tile_overlap_factor_height or self.tile_overlap_factor_height
        self.tile_overlap_factor_width = tile_overlap_factor_width or self.tile_overlap_factor_width

    def disable_tiling(self) -> None:
        self.use_tiling = False

    def enable_slicing(self) -> None:
        self.use_slicing = True

    def disable_slicing(self) -> None:
        self.use_slicing = False

    def invalidate_engine(self):
        object.__setattr__(self, "_b200_engine", None)
        object.__setattr__(self, "_b200_enc_engine", None)

    def load_state_dict(self, state_dict, strict: bool = True, **kw):
        """A full checkpoint loads strictly, like the reference class.  A dict that holds ONE half only (`decoder.*` or
        `encoder.*` — decode-only deployments, the test fixtures) loads that half strictly and leaves the other half untouched."""
        halves = {k.split(".", 1)[0] for k in state_dict}
        if strict and halves and halves < {"encoder", "decoder"}:
            res = super().load_state_dict(state_dict, strict=False, **kw)
            other = ({"encoder", "decoder"} - halves).pop() + "."
            missing = [k for k in res.missing_keys if not k.startswith(other)]
            if missing or res.unexpected_keys:
                raise RuntimeError(f"Error(s) in loading state_dict for AutoencoderKLCogVideoX: missing {missing[:5]}, "
                                   f"unexpected {list(res.unexpected_keys)[:5]}")
            self.invalidate_engine()
            return res
        res = super().load_state_dict(state_dict, strict=strict, **kw)
        self.invalidate_engine()
        return res

    def forward(self, *a, **k):
        raise RuntimeError("call decode(); AutoencoderKLCogVideoX.forward (encode + decode) is not part of the denoising path")


def attach_vae(vae: nn.Module) -> nn.Module:
    """Bind the B200 decoder engine to an already constructed stock diffusers `AutoencoderKLCogVideoX` (parameters read in
    place through its state dict): replaces decode / _decode / tiled_decode / blend_v / blend_h; tiling and slicing flags,
    tile sizes and overlap factors keep coming from the object itself."""
    for name in ("_engine", "_decode_tile", "_decode", "decode", "tiled_decode", "blend_v", "blend_h"):
        object.__setattr__(vae, name, types.MethodType(getattr(_DecodeMixin, name), vae))
    for name in ("_enc_engine", "_encode_image", "_encode", "encode", "tiled_encode"):
        object.__setattr__(vae, name, types.MethodType(getattr(_EncodeMixin, name), vae))
    object.__setattr__(vae, "_blend", _DecodeMixin._blend)
    if getattr(vae, "post_quant_conv", None) is not None or getattr(vae, "quant_conv", None) is not None:
        raise NotImplementedError("quant_conv / post_quant_conv are not used by CogVideoX VAEs and are not implemented")
    return vae
