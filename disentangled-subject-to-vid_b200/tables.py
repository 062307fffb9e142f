"""Host-side position tables of the hot path (computed once per pipeline call, then resident on the device).

Index / table arithmetic must be bit-exact with the reference, so the same numpy / torch primitives are used:
  - 3-D RoPE cos/sin tables: D/models/embeddings.py:505-570 (get_3d_rotary_pos_embed), :673-736 (1-D frequencies),
    crop region D/pipelines/cogvideo/pipeline_cogvideox.py:62-77, table sizing / slicing :436-460 and
    S/custom_cogvideox_pipe.py:223-235 (temporal size F+1: position 0 = reference image, 1..F = video frames);
  - 3-D sincos positional embedding of CogVideoX-2B: D/models/embeddings.py:81-125.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np
import torch


def crop_region_for_grid(grid_hw: Tuple[int, int], target_w: int, target_h: int):
    h, w = grid_hw
    if h / w > target_h / target_w:
        new_h, new_w = target_h, int(round(target_h / h * w))
    else:
        new_w, new_h = target_w, int(round(target_w / w * h))
    top = int(round((target_h - new_h) / 2.0))
    left = int(round((target_w - new_w) / 2.0))
    return (top, left), (top + new_h, left + new_w)


def _axis_freqs(dim: int, positions: np.ndarray, theta: float = 10000.0):
    inv = 1.0 / (theta ** (torch.arange(0, dim, 2, dtype=torch.float32)[: dim // 2] / dim))
    ang = torch.outer(torch.from_numpy(positions), inv)
    return ang.cos().repeat_interleave(2, dim=1).float(), ang.sin().repeat_interleave(2, dim=1).float()


def rope_table_3d(head_dim: int, crop, grid_hw: Tuple[int, int], frames: int):
    """cos, sin fp32 [frames*gh*gw, head_dim]; channel split t/h/w = d/4, 3d/8, 3d/8; frame-major row order."""
    (top, left), (bottom, right) = crop
    gh, gw = grid_hw
    pos_h = np.linspace(top, bottom, gh, endpoint=False, dtype=np.float32)
    pos_w = np.linspace(left, right, gw, endpoint=False, dtype=np.float32)
    pos_t = np.linspace(0, frames, frames, endpoint=False, dtype=np.float32)
    ct, st = _axis_freqs(head_dim // 4, pos_t)
    ch, sh = _axis_freqs(head_dim // 8 * 3, pos_h)
    cw, sw = _axis_freqs(head_dim // 8 * 3, pos_w)

    def lattice(a_t, a_h, a_w):
        parts = [a_t[:, None, None, :].expand(frames, gh, gw, -1), a_h[None, :, None, :].expand(frames, gh, gw, -1),
                 a_w[None, None, :, :].expand(frames, gh, gw, -1)]
        return torch.cat(parts, dim=-1).reshape(frames * gh * gw, -1).contiguous()

    return lattice(ct, ch, cw), lattice(st, sh, sw)


def joint_rope_table(height: int, width: int, latent_frames: int, head_dim: int = 64, patch: int = 2, vae_scale: int = 8):
    """The table the custom pipeline builds (temporal size latent_frames+1).  Rows [0, n) belong to the reference image,
    rows [n, (F+1)n) to the video — exactly the [ref | video] row order of the joint token buffer, so ONE table serves
    both RoPE applications of the attention processor.  The reference hard-codes 14 and 1350 (480x720, <=13 frames);
    generalised to latent_frames+1 and (H/16)(W/16)."""
    gh, gw = height // (vae_scale * patch), width // (vae_scale * patch)
    crop = crop_region_for_grid((gh, gw), 720 // (vae_scale * patch), 480 // (vae_scale * patch))
    return rope_table_3d(head_dim, crop, (gh, gw), latent_frames + 1)


def _sincos_axis(dim: int, pos: np.ndarray) -> np.ndarray:
    omega = np.arange(dim // 2, dtype=np.float64)
    omega /= dim / 2.0
    omega = 1.0 / 10000**omega
    ang = np.einsum("m,d->md", pos.reshape(-1), omega)
    return np.concatenate([np.sin(ang), np.cos(ang)], axis=1)


def sincos_table_3d(embed_dim: int, grid_w: int, grid_h: int, frames: int, spatial_scale: float, temporal_scale: float):
    """fp32 [frames*grid_h*grid_w, embed_dim] (temporal quarter first, then spatial w-half, h-half)."""
    if embed_dim % 4:
        raise ValueError("`embed_dim` must be divisible by 4")
    d_sp, d_t = 3 * embed_dim // 4, embed_dim // 4
    gh = np.arange(grid_h, dtype=np.float32) / spatial_scale
    gw = np.arange(grid_w, dtype=np.float32) / spatial_scale
    mesh = np.stack(np.meshgrid(gw, gh), axis=0).reshape(2, 1, grid_h, grid_w)
    spatial = np.concatenate([_sincos_axis(d_sp // 2, mesh[0]), _sincos_axis(d_sp // 2, mesh[1])], axis=1)
    temporal = _sincos_axis(d_t, np.arange(frames, dtype=np.float32) / temporal_scale)
    spatial = np.repeat(spatial[np.newaxis], frames, axis=0)
    temporal = np.repeat(temporal[:, np.newaxis], grid_h * grid_w, axis=1)
    table = np.concatenate([temporal, spatial], axis=-1)
    return torch.from_numpy(table).flatten(0, 1).to(torch.float32)
