"""CPU oracle for row V of SURVEY §8: the CogVideoX 3D causal VAE *decoder* and its frame-batched / tiled `decode`.

TEST INFRASTRUCTURE ONLY (same rules as s2v_oracle.py): imported by tests/, smoke() and bench.py's CPU arm, never by the
product.  Plain torch-CPU functional restatement; every function cites the reference lines it follows
(A/ = /root/reference/diffusers/src/diffusers/models/autoencoders/autoencoder_kl_cogvideox.py,
 U/ = /root/reference/diffusers/src/diffusers/models/upsampling.py).  Pinned against the reference's own
`AutoencoderKLCogVideoX` run in the build container (tests/golden/make_golden.py:gen_vae -> tests/golden/vae_tiny.pt).

Parameters: flat dict keyed exactly like the reference `state_dict()` ("decoder.conv_in.conv.weight",
"decoder.up_blocks.0.resnets.1.norm1.conv_y.conv.bias", "decoder.up_blocks.0.upsamplers.0.conv.weight", ...).
Tensors are [B, C, T, H, W] like the reference.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Params = Dict[str, torch.Tensor]
Cache = Dict[str, torch.Tensor]


@dataclass
class VaeConfig:
    """`register_to_config` fields of AutoencoderKLCogVideoX used by decode (A/:1019-1052) and the tiling constants
    derived from them (A/:1100-1115)."""

    latent_channels: int = 16
    out_channels: int = 3
    block_out_channels: Tuple[int, ...] = (128, 256, 256, 512)
    layers_per_block: int = 3
    norm_eps: float = 1e-6
    norm_num_groups: int = 32
    temporal_compression_ratio: float = 4
    sample_height: int = 480
    sample_width: int = 720
    scaling_factor: float = 0.7
    num_latent_frames_batch_size: int = 2
    tile_overlap_factor_height: float = 1 / 6
    tile_overlap_factor_width: float = 1 / 5

    @property
    def tile_sample_min_height(self) -> int:
        return self.sample_height // 2

    @property
    def tile_sample_min_width(self) -> int:
        return self.sample_width // 2

    @property
    def tile_latent_min_height(self) -> int:
        return int(self.tile_sample_min_height / (2 ** (len(self.block_out_channels) - 1)))

    @property
    def tile_latent_min_width(self) -> int:
        return int(self.tile_sample_min_width / (2 ** (len(self.block_out_channels) - 1)))


# --------------------------------------------------------------------------- parameter synthesis (tests / bench)
def decoder_param_shapes(cfg: VaeConfig) -> Dict[str, Tuple[int, ...]]:
    """Names and shapes of the decoder parameters (CogVideoXDecoder3D.__init__, A/:842-913)."""
    ch = list(reversed(cfg.block_out_channels))
    z = cfg.latent_channels
    shapes: Dict[str, Tuple[int, ...]] = {}

    def conv3(name, cin, cout, k):
        shapes[f"{name}.conv.weight"] = (cout, cin, k, k, k)
        shapes[f"{name}.conv.bias"] = (cout,)

    def spatial_norm(name, c):
        shapes[f"{name}.norm_layer.weight"] = (c,)
        shapes[f"{name}.norm_layer.bias"] = (c,)
        conv3(f"{name}.conv_y", z, c, 1)
        conv3(f"{name}.conv_b", z, c, 1)

    def resnet(name, cin, cout):
        spatial_norm(f"{name}.norm1", cin)
        conv3(f"{name}.conv1", cin, cout, 3)
        spatial_norm(f"{name}.norm2", cout)
        conv3(f"{name}.conv2", cout, cout, 3)
        if cin != cout:  # conv_shortcut=False default -> CogVideoXSafeConv3d 1x1x1 (A/:261-270)
            shapes[f"{name}.conv_shortcut.weight"] = (cout, cin, 1, 1, 1)
            shapes[f"{name}.conv_shortcut.bias"] = (cout,)

    conv3("decoder.conv_in", z, ch[0], 3)
    for i in range(2):
        resnet(f"decoder.mid_block.resnets.{i}", ch[0], ch[0])
    prev = ch[0]
    for b, c in enumerate(ch):
        for i in range(cfg.layers_per_block + 1):
            resnet(f"decoder.up_blocks.{b}.resnets.{i}", prev if i == 0 else c, c)
        if b != len(ch) - 1:
            shapes[f"decoder.up_blocks.{b}.upsamplers.0.conv.weight"] = (c, c, 3, 3)
            shapes[f"decoder.up_blocks.{b}.upsamplers.0.conv.bias"] = (c,)
        prev = c
    spatial_norm("decoder.norm_out", ch[-1])
    conv3("decoder.conv_out", ch[-1], cfg.out_channels, 3)
    return shapes


def synth_decoder_params(cfg: VaeConfig, seed: int = 0) -> Params:
    """Deterministic non-trivial weights: fan-in scaled normals for conv weights (so activations stay O(1) through
    ~30 layers), normal(0, 0.05) biases, GroupNorm weight 1 + normal(0, 0.1), bias normal(0, 0.1)."""
    g = torch.Generator().manual_seed(seed)
    p: Params = {}
    for name, shape in decoder_param_shapes(cfg).items():
        if "norm_layer.weight" in name:
            p[name] = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif "norm_layer.bias" in name:
            p[name] = 0.1 * torch.randn(shape, generator=g)
        elif name.endswith("bias"):
            p[name] = 0.05 * torch.randn(shape, generator=g)
        else:
            fan_in = int(np.prod(shape[1:]))
            p[name] = torch.randn(shape, generator=g) / (fan_in ** 0.5)
    return p


# --------------------------------------------------------------------------- layers
def causal_conv3d(p: Params, name: str, x: torch.Tensor, cache: Optional[torch.Tensor]) -> Tuple[torch.Tensor, torch.Tensor]:
    """CogVideoXCausalConv3d.forward (A/:129-137): temporal left context = the cache (last k-1 input frames of the
    previous call) or k-1 copies of the first frame (A/:120-127); zero padding in H and W; stride/dilation 1.
    (CogVideoXSafeConv3d's >2 GB chunking, A/:43-66, only splits the same convolution along T.)"""
    w, b = p[f"{name}.conv.weight"], p[f"{name}.conv.bias"]
    k = w.shape[2]
    if k > 1:
        ctx = [cache] if cache is not None else [x[:, :, :1]] * (k - 1)
        x = torch.cat(ctx + [x], dim=2)
    new_cache = x[:, :, -k + 1:].clone() if k > 1 else x[:, :, :0].clone()
    pad = w.shape[3] // 2
    x = F.pad(x, (pad, pad, pad, pad), mode="constant", value=0)
    return F.conv3d(x, w, b), new_cache


def spatial_norm(p: Params, name: str, f: torch.Tensor, zq: torch.Tensor, groups: int) -> torch.Tensor:
    """CogVideoXSpatialNorm3D.forward (A/:167-188): GroupNorm(f; eps 1e-6) * conv_y(zq') + conv_b(zq'), zq' = zq
    nearest-resized to f's (T, H, W) with the first frame handled separately when T is odd and > 1.
    conv_y / conv_b are 1x1x1 (no temporal context, so their caches are empty)."""
    if f.shape[2] > 1 and f.shape[2] % 2 == 1:
        z_first = F.interpolate(zq[:, :, :1], size=(1,) + tuple(f.shape[-2:]))
        z_rest = F.interpolate(zq[:, :, 1:], size=(f.shape[2] - 1,) + tuple(f.shape[-2:]))
        zq = torch.cat([z_first, z_rest], dim=2)
    else:
        zq = F.interpolate(zq, size=tuple(f.shape[-3:]))
    y, _ = causal_conv3d(p, f"{name}.conv_y", zq, None)
    b, _ = causal_conv3d(p, f"{name}.conv_b", zq, None)
    norm_f = F.group_norm(f, groups, p[f"{name}.norm_layer.weight"], p[f"{name}.norm_layer.bias"], eps=1e-6)
    return norm_f * y + b


def resnet_block(p: Params, name: str, x: torch.Tensor, zq: torch.Tensor, groups: int, cache: Cache, new_cache: Cache) -> torch.Tensor:
    """CogVideoXResnetBlock3D.forward with zq and temb=None (A/:278-319): norm1 -> SiLU -> conv1 -> norm2 -> SiLU ->
    conv2, plus the (1x1x1 conv) shortcut when the channel count changes."""
    h = F.silu(spatial_norm(p, f"{name}.norm1", x, zq, groups))
    h, new_cache[f"{name}.conv1"] = causal_conv3d(p, f"{name}.conv1", h, cache.get(f"{name}.conv1"))
    h = F.silu(spatial_norm(p, f"{name}.norm2", h, zq, groups))
    h, new_cache[f"{name}.conv2"] = causal_conv3d(p, f"{name}.conv2", h, cache.get(f"{name}.conv2"))
    if f"{name}.conv_shortcut.weight" in p:
        x = F.conv3d(x, p[f"{name}.conv_shortcut.weight"], p[f"{name}.conv_shortcut.bias"])
    return h + x


def upsample3d(p: Params, name: str, x: torch.Tensor, compress_time: bool) -> torch.Tensor:
    """CogVideoXUpsample3D.forward (U/:384-412): nearest 2x in H, W (and in T when compress_time, keeping the first
    frame single when T is odd and > 1), then a per-frame 3x3 Conv2d."""
    if compress_time:
        if x.shape[2] > 1 and x.shape[2] % 2 == 1:
            first = F.interpolate(x[:, :, 0], scale_factor=2.0)[:, :, None]
            rest = F.interpolate(x[:, :, 1:], scale_factor=2.0)
            x = torch.cat([first, rest], dim=2)
        elif x.shape[2] > 1:
            x = F.interpolate(x, scale_factor=2.0)
        else:
            x = F.interpolate(x.squeeze(2), scale_factor=2.0)[:, :, None]
    else:
        b, c, t, h, w = x.shape
        x = F.interpolate(x.permute(0, 2, 1, 3, 4).reshape(b * t, c, h, w), scale_factor=2.0)
        x = x.reshape(b, t, c, *x.shape[2:]).permute(0, 2, 1, 3, 4)
    b, c, t, h, w = x.shape
    y = F.conv2d(x.permute(0, 2, 1, 3, 4).reshape(b * t, c, h, w), p[f"{name}.conv.weight"], p[f"{name}.conv.bias"], padding=1)
    return y.reshape(b, t, *y.shape[1:]).permute(0, 2, 1, 3, 4)


def decoder_forward(p: Params, cfg: VaeConfig, sample: torch.Tensor, cache: Optional[Cache]) -> Tuple[torch.Tensor, Cache]:
    """CogVideoXDecoder3D.forward (A/:921-981): conv_in -> mid block (2 resnets) -> up blocks (layers_per_block+1 resnets
    each, upsample in all but the last, temporal upsample in the first log2(ratio)) -> SpatialNorm -> SiLU -> conv_out.
    `sample` doubles as zq for every SpatialNorm."""
    cache = cache or {}
    new_cache: Cache = {}
    g = cfg.norm_num_groups
    h, new_cache["decoder.conv_in"] = causal_conv3d(p, "decoder.conv_in", sample, cache.get("decoder.conv_in"))
    for i in range(2):
        h = resnet_block(p, f"decoder.mid_block.resnets.{i}", h, sample, g, cache, new_cache)
    n_blocks = len(cfg.block_out_channels)
    t_levels = int(np.log2(cfg.temporal_compression_ratio))
    for b in range(n_blocks):
        for i in range(cfg.layers_per_block + 1):
            h = resnet_block(p, f"decoder.up_blocks.{b}.resnets.{i}", h, sample, g, cache, new_cache)
        if b != n_blocks - 1:
            h = upsample3d(p, f"decoder.up_blocks.{b}.upsamplers.0", h, compress_time=b < t_levels)
    h = F.silu(spatial_norm(p, "decoder.norm_out", h, sample, g))
    h, new_cache["decoder.conv_out"] = causal_conv3d(p, "decoder.conv_out", h, cache.get("decoder.conv_out"))
    return h, new_cache


# --------------------------------------------------------------------------- decode (frame batching, tiling, blending)
def frame_batches(num_frames: int, batch: int) -> List[Tuple[int, int]]:
    """Temporal batching of `_decode` / `tiled_decode` (A/:1238-1247, :1414-1419): max(T // batch, 1) calls; the first
    one also takes the T % batch remainder.  13 latent frames, batch 2 -> [0:3] [3:5] ... [11:13]."""
    n = max(num_frames // batch, 1)
    rem = num_frames % batch
    return [(batch * k + (0 if k == 0 else rem), batch * (k + 1) + rem) for k in range(n)]


def decode_untiled(p: Params, cfg: VaeConfig, z: torch.Tensor) -> torch.Tensor:
    """`_decode` without tiling (A/:1238-1252): conv caches carried across the temporal batches."""
    cache = None
    out = []
    for s, e in frame_batches(z.shape[2], cfg.num_latent_frames_batch_size):
        y, cache = decoder_forward(p, cfg, z[:, :, s:e], cache)
        out.append(y)
    return torch.cat(out, dim=2)


def blend_v(a: torch.Tensor, b: torch.Tensor, extent: int) -> torch.Tensor:
    """blend_v (A/:1284-1290): rows [0, extent) of b become a linear ramp from a's last `extent` rows."""
    extent = min(a.shape[3], b.shape[3], extent)
    for y in range(extent):
        b[:, :, :, y, :] = a[:, :, :, -extent + y, :] * (1 - y / extent) + b[:, :, :, y, :] * (y / extent)
    return b


def blend_h(a: torch.Tensor, b: torch.Tensor, extent: int) -> torch.Tensor:
    """blend_h (A/:1292-1298)."""
    extent = min(a.shape[4], b.shape[4], extent)
    for x in range(extent):
        b[:, :, :, :, x] = a[:, :, :, :, -extent + x] * (1 - x / extent) + b[:, :, :, :, x] * (x / extent)
    return b


def tile_plan(cfg: VaeConfig, height: int, width: int):
    """Tile origins, blend extents and crop limits of tiled_decode (A/:1398-1404)."""
    oh = int(cfg.tile_latent_min_height * (1 - cfg.tile_overlap_factor_height))
    ow = int(cfg.tile_latent_min_width * (1 - cfg.tile_overlap_factor_width))
    bh = int(cfg.tile_sample_min_height * cfg.tile_overlap_factor_height)
    bw = int(cfg.tile_sample_min_width * cfg.tile_overlap_factor_width)
    return dict(rows=list(range(0, height, oh)), cols=list(range(0, width, ow)), blend_h=bh, blend_w=bw,
                limit_h=cfg.tile_sample_min_height - bh, limit_w=cfg.tile_sample_min_width - bw)


def tiled_decode(p: Params, cfg: VaeConfig, z: torch.Tensor) -> torch.Tensor:
    """tiled_decode (A/:1374-1455): overlapping latent tiles, each decoded with its own conv-cache chain over the
    temporal batches; each tile is blended with its upper and left neighbours (NB: the neighbours were themselves
    blended in place earlier — blend_* mutate `b` — and that order is reproduced) and cropped."""
    tp = tile_plan(cfg, z.shape[3], z.shape[4])
    rows = []
    for i in tp["rows"]:
        row = []
        for j in tp["cols"]:
            cache = None
            time = []
            for s, e in frame_batches(z.shape[2], cfg.num_latent_frames_batch_size):
                tile = z[:, :, s:e, i:i + cfg.tile_latent_min_height, j:j + cfg.tile_latent_min_width]
                y, cache = decoder_forward(p, cfg, tile, cache)
                time.append(y)
            row.append(torch.cat(time, dim=2))
        rows.append(row)
    result_rows = []
    for i, row in enumerate(rows):
        result_row = []
        for j, tile in enumerate(row):
            if i > 0:
                tile = blend_v(rows[i - 1][j], tile, tp["blend_h"])
            if j > 0:
                tile = blend_h(row[j - 1], tile, tp["blend_w"])
            result_row.append(tile[:, :, :, :tp["limit_h"], :tp["limit_w"]])
        result_rows.append(torch.cat(result_row, dim=4))
    return torch.cat(result_rows, dim=3)


def decode(p: Params, cfg: VaeConfig, z: torch.Tensor, use_tiling: bool = True) -> torch.Tensor:
    """AutoencoderKLCogVideoX.decode / _decode (A/:1231-1282): per-sample slicing is a pure batch split; tiling applies
    when the latent is larger than the minimum tile (A/:1234-1235)."""
    outs = []
    for zs in z.split(1):
        if use_tiling and (zs.shape[4] > cfg.tile_latent_min_width or zs.shape[3] > cfg.tile_latent_min_height):
            outs.append(tiled_decode(p, cfg, zs))
        else:
            outs.append(decode_untiled(p, cfg, zs))
    return torch.cat(outs)


def decode_latents(p: Params, cfg: VaeConfig, latents: torch.Tensor, use_tiling: bool = True) -> torch.Tensor:
    """CogVideoXPipeline.decode_latents (D/pipelines/cogvideo/pipeline_cogvideox.py:346-351): [B, F, C, H, W] latents ->
    permute -> 1/scaling_factor -> vae.decode."""
    z = latents.permute(0, 2, 1, 3, 4)
    z = 1 / cfg.scaling_factor * z
    return decode(p, cfg, z, use_tiling)


# ----------------------------------------------------------------------------------------------- host glue after the decoder
def frames_uint8(video: torch.Tensor, round_half_even: bool = False):
    """uint8 frames [B,F,H,W,C] the reference derives from the decoder output `video` [B,C,F,H,W] (any float dtype):
    D/video_processor.py:89-113 postprocess_video -> D/image_processor.py:227-239 denormalize `(x / 2 + 0.5).clamp(0, 1)`
    (two tensor ops in the video's dtype, so bf16 rounds twice) -> :196-208 pt_to_numpy (`.float().numpy()`), then either
    D/utils/export_utils.py:177-178 `(frame * 255).astype(np.uint8)` (truncation; the path S/video_generate.py:74-75 takes)
    or D/image_processor.py:133-150 numpy_to_pil `(images * 255).round().astype("uint8")` (round half to even)."""
    import numpy as np

    outs = []
    for b in range(video.shape[0]):
        frames = video[b].permute(1, 0, 2, 3)                       # [F,C,H,W]
        frames = (frames / 2 + 0.5).clamp(0, 1)
        arr = frames.cpu().permute(0, 2, 3, 1).float().numpy()      # [F,H,W,C] fp32
        arr = arr * 255
        outs.append((arr.round() if round_half_even else arr).astype(np.uint8))
    return np.stack(outs)


# =============================================================================================== encoder
# CogVideoXEncoder3D (A/:658-814) for the reference-image path of S/video_generate.py:26-38: ONE frame per call, so the
# temporal compression of CogVideoXDownsample3D (D/models/downsampling.py:322-343) is the identity and every causal conv sees
# its single frame three times.  Multi-frame (video) encoding is outside the subject-to-video inference path.
def encoder_param_shapes(cfg: VaeConfig, in_channels: int = 3) -> Dict[str, Tuple[int, ...]]:
    """Names and shapes of the encoder parameters (CogVideoXEncoder3D.__init__, A/:682-753)."""
    ch = list(cfg.block_out_channels)
    shapes: Dict[str, Tuple[int, ...]] = {}

    def conv3(name, cin, cout, k):
        shapes[f"{name}.conv.weight"] = (cout, cin, k, k, k)
        shapes[f"{name}.conv.bias"] = (cout,)

    def resnet(name, cin, cout):   # spatial_norm_dim=None -> plain nn.GroupNorm (A/:241-243)
        shapes[f"{name}.norm1.weight"] = (cin,)
        shapes[f"{name}.norm1.bias"] = (cin,)
        conv3(f"{name}.conv1", cin, cout, 3)
        shapes[f"{name}.norm2.weight"] = (cout,)
        shapes[f"{name}.norm2.bias"] = (cout,)
        conv3(f"{name}.conv2", cout, cout, 3)
        if cin != cout:
            shapes[f"{name}.conv_shortcut.weight"] = (cout, cin, 1, 1, 1)
            shapes[f"{name}.conv_shortcut.bias"] = (cout,)

    conv3("encoder.conv_in", in_channels, ch[0], 3)
    prev = ch[0]
    for b, c in enumerate(ch):
        for i in range(cfg.layers_per_block):
            resnet(f"encoder.down_blocks.{b}.resnets.{i}", prev if i == 0 else c, c)
        if b != len(ch) - 1:
            shapes[f"encoder.down_blocks.{b}.downsamplers.0.conv.weight"] = (c, c, 3, 3)
            shapes[f"encoder.down_blocks.{b}.downsamplers.0.conv.bias"] = (c,)
        prev = c
    for i in range(2):
        resnet(f"encoder.mid_block.resnets.{i}", ch[-1], ch[-1])
    shapes["encoder.norm_out.weight"] = (ch[-1],)
    shapes["encoder.norm_out.bias"] = (ch[-1],)
    conv3("encoder.conv_out", ch[-1], 2 * cfg.latent_channels, 3)
    return shapes


def synth_encoder_params(cfg: VaeConfig, seed: int = 0) -> Params:
    g = torch.Generator().manual_seed(seed)
    p: Params = {}
    for name, shape in encoder_param_shapes(cfg).items():
        if "norm" in name and name.endswith("weight") and len(shape) == 1:
            p[name] = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif "norm" in name and name.endswith("bias"):
            p[name] = 0.1 * torch.randn(shape, generator=g)
        elif name.endswith("bias"):
            p[name] = 0.05 * torch.randn(shape, generator=g)
        else:
            p[name] = torch.randn(shape, generator=g) / (int(np.prod(shape[1:])) ** 0.5)
    return p


def resnet_block_plain(p: Params, name: str, x: torch.Tensor, groups: int, eps: float) -> torch.Tensor:
    """CogVideoXResnetBlock3D.forward with zq=None, temb=None (A/:278-319), no conv cache (single call)."""
    h = F.silu(F.group_norm(x, groups, p[f"{name}.norm1.weight"], p[f"{name}.norm1.bias"], eps))
    h, _ = causal_conv3d(p, f"{name}.conv1", h, None)
    h = F.silu(F.group_norm(h, groups, p[f"{name}.norm2.weight"], p[f"{name}.norm2.bias"], eps))
    h, _ = causal_conv3d(p, f"{name}.conv2", h, None)
    if f"{name}.conv_shortcut.weight" in p:
        x = F.conv3d(x, p[f"{name}.conv_shortcut.weight"], p[f"{name}.conv_shortcut.bias"])
    return h + x


def downsample3d_single_frame(p: Params, name: str, x: torch.Tensor) -> torch.Tensor:
    """CogVideoXDownsample3D.forward (D/models/downsampling.py:322-353) for ONE frame: with an odd frame count the first frame
    is kept and `x_rest` is empty (:329-335), so compress_time changes nothing; then F.pad (0,1,0,1) and the stride-2 3x3
    Conv2d applied per frame."""
    b, c, t, h, w = x.shape
    assert t == 1, "the oracle covers the single-frame (reference image) path"
    y = F.pad(x, (0, 1, 0, 1), mode="constant", value=0)
    y = y.permute(0, 2, 1, 3, 4).reshape(b * t, c, h + 1, w + 1)
    y = F.conv2d(y, p[f"{name}.conv.weight"], p[f"{name}.conv.bias"], stride=2)
    return y.reshape(b, t, y.shape[1], y.shape[2], y.shape[3]).permute(0, 2, 1, 3, 4)


def encoder_forward(p: Params, cfg: VaeConfig, sample: torch.Tensor) -> torch.Tensor:
    """CogVideoXEncoder3D.forward (A/:755-814) -> the 2*latent_channels moments (mean | logvar)."""
    G, eps = cfg.norm_num_groups, cfg.norm_eps
    h, _ = causal_conv3d(p, "encoder.conv_in", sample, None)
    n = len(cfg.block_out_channels)
    for b in range(n):
        for i in range(cfg.layers_per_block):
            h = resnet_block_plain(p, f"encoder.down_blocks.{b}.resnets.{i}", h, G, eps)
        if b != n - 1:
            h = downsample3d_single_frame(p, f"encoder.down_blocks.{b}.downsamplers.0", h)
    for i in range(2):
        h = resnet_block_plain(p, f"encoder.mid_block.resnets.{i}", h, G, eps)
    h = F.silu(F.group_norm(h, G, p["encoder.norm_out.weight"], p["encoder.norm_out.bias"], 1e-6))   # A/:751 eps is literal 1e-6
    h, _ = causal_conv3d(p, "encoder.conv_out", h, None)
    return h


def tiled_encode(p: Params, cfg: VaeConfig, x: torch.Tensor) -> torch.Tensor:
    """AutoencoderKLCogVideoX.tiled_encode (A/:1300-1372) for one frame: overlapping sample tiles, each through the encoder,
    blended over the latent overlap and cropped to the row / column limits."""
    oh = int(cfg.tile_sample_min_height * (1 - cfg.tile_overlap_factor_height))
    ow = int(cfg.tile_sample_min_width * (1 - cfg.tile_overlap_factor_width))
    bh = int(cfg.tile_latent_min_height * cfg.tile_overlap_factor_height)
    bw = int(cfg.tile_latent_min_width * cfg.tile_overlap_factor_width)
    lh, lw = cfg.tile_latent_min_height - bh, cfg.tile_latent_min_width - bw
    H, W = x.shape[3], x.shape[4]
    rows = []
    for i in range(0, H, oh):
        rows.append([encoder_forward(p, cfg, x[:, :, :, i:i + cfg.tile_sample_min_height, j:j + cfg.tile_sample_min_width])
                     for j in range(0, W, ow)])
    out_rows = []
    for i, row in enumerate(rows):
        out = []
        for j, tile in enumerate(row):
            if i > 0:
                tile = blend_v(rows[i - 1][j], tile, bh)
            if j > 0:
                tile = blend_h(row[j - 1], tile, bw)
            out.append(tile[:, :, :, :lh, :lw])
        out_rows.append(torch.cat(out, dim=4))
    return torch.cat(out_rows, dim=3)


def encode_moments(p: Params, cfg: VaeConfig, x: torch.Tensor, use_tiling: bool = True) -> torch.Tensor:
    """AutoencoderKLCogVideoX.encode/_encode (A/:1177-1229) for x [B, 3, 1, H, W]: the moments tensor the reference wraps in
    DiagonalGaussianDistribution (quant_conv is None for CogVideoX)."""
    assert x.shape[2] == 1
    outs = []
    for xs in x.split(1):   # use_slicing
        if use_tiling and (x.shape[4] > cfg.tile_sample_min_width or x.shape[3] > cfg.tile_sample_min_height):
            outs.append(tiled_encode(p, cfg, xs))
        else:
            outs.append(encoder_forward(p, cfg, xs))
    return torch.cat(outs)


def gaussian_sample(moments: torch.Tensor, noise: torch.Tensor) -> torch.Tensor:
    """DiagonalGaussianDistribution (D/models/autoencoders/vae.py:691-744): mean, logvar = chunk(2, dim=1); logvar clamped to
    [-30, 20]; std = exp(0.5 logvar); sample = mean + std * noise."""
    mean, logvar = torch.chunk(moments, 2, dim=1)
    logvar = torch.clamp(logvar, -30.0, 20.0)
    return mean + torch.exp(0.5 * logvar) * noise


def reference_image_latents(p: Params, cfg: VaeConfig, image_u8: "np.ndarray", noise: torch.Tensor, use_tiling: bool = True,
                            dtype=torch.float32) -> torch.Tensor:
    """S/video_generate.py:26-38: uint8 RGB [H, W, 3] -> float / 255 * 2 - 1 -> [1, 3, 1, H, W] -> vae.encode -> sample *
    scaling_factor -> [1, 1, C, h, w] (the `ref_img_states` pipeline argument)."""
    x = torch.from_numpy(np.expand_dims(image_u8, 0)).float() / 255.0 * 2.0 - 1.0
    x = x.permute(0, 3, 1, 2).unsqueeze(0).permute(0, 2, 1, 3, 4).to(dtype)
    lat = gaussian_sample(encode_moments(p, cfg, x, use_tiling), noise) * cfg.scaling_factor
    return lat.permute(0, 2, 1, 3, 4)
