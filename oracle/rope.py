"""ORACLE (test infrastructure): numpy/torch restatement of the reference's 3D-RoPE tables and position grids.

Follows longvgen/models/embeddings.py:571-707 (get_3d_rotary_pos_embed[_v2]), :774-837 (get_1d_rotary_pos_embed) and
the grid builders of longvgen/pipeline/pipeline_cogvideox_mp_fifo.py:769-813,1061-1103.  Pinned against the reference
by tests/golden/rope_*.pt (bit-exact class for the position grids; table values are compared exactly as well because
both sides call the same torch fp32 cos/sin).
"""
from __future__ import annotations

from typing import Tuple

import numpy as np
import torch


def rope_1d(dim: int, pos: np.ndarray, theta: float = 10000.0) -> Tuple[torch.Tensor, torch.Tensor]:
    """embeddings.py:812-825: freqs = 1/theta^(2i/dim) in fp32, outer(pos, freqs), cos/sin repeat-interleaved by 2."""
    pos_t = torch.from_numpy(np.asarray(pos))
    freqs = 1.0 / (theta ** (torch.arange(0, dim, 2, dtype=torch.float32)[: dim // 2] / dim))
    ang = torch.outer(pos_t, freqs)
    return ang.cos().repeat_interleave(2, dim=1).float(), ang.sin().repeat_interleave(2, dim=1).float()


def rope_3d_from_grids(head_dim: int, grid_t, grid_h, grid_w, dim_t=None, dim_h=None, dim_w=None):
    """embeddings.py:641-707: per-axis tables concatenated as [t | h | w] along the feature axis, row-major over (t,h,w)."""
    dim_t = head_dim // 4 if dim_t is None else dim_t
    dim_h = head_dim // 8 * 3 if dim_h is None else dim_h
    dim_w = head_dim // 8 * 3 if dim_w is None else dim_w
    T, H, W = len(grid_t), len(grid_h), len(grid_w)
    out = []
    for which in (0, 1):
        t = rope_1d(dim_t, grid_t)[which][:, None, None, :].expand(T, H, W, dim_t)
        h = rope_1d(dim_h, grid_h)[which][None, :, None, :].expand(T, H, W, dim_h)
        w = rope_1d(dim_w, grid_w)[which][None, None, :, :].expand(T, H, W, dim_w)
        out.append(torch.cat([t, h, w], dim=-1).reshape(T * H * W, -1))
    return out[0], out[1]


def rope_3d(head_dim: int, crops_coords, grid_size):
    """embeddings.py:571-639: grids are fp32 linspace(start, stop, n, endpoint=False) per axis."""
    (t0, h0, w0), (t1, h1, w1) = crops_coords
    T, H, W = grid_size
    gt = np.linspace(t0, t1, T, endpoint=False, dtype=np.float32)
    gh = np.linspace(h0, h1, H, endpoint=False, dtype=np.float32)
    gw = np.linspace(w0, w1, W, endpoint=False, dtype=np.float32)
    return rope_3d_from_grids(head_dim, gt, gh, gw)


def window_rope(head_dim: int, frames: int, grid_h: int, grid_w: int):
    """pipeline_cogvideox_mp_fifo.py:769-795 at the native 480x720 resolution: crop region == full grid."""
    return rope_3d(head_dim, [[0, 0, 0], [frames, grid_h, grid_w]], (frames, grid_h, grid_w))


def vip_grids(latent_h: int, latent_w: int, patch: int, num_chunks: int, frames_per_chunk: int, vip_frames_per_chunk: int,
              h_queries: int, w_queries: int, start_frame_idx: float):
    """pipeline_cogvideox_mp_fifo.py:1061-1103: position grids of the video tokens (image grid) and of the condensed
    tokens (condition grid, temporal positions start_frame_idx + chunk*frames + k*frames/vip_frames)."""
    gh, gw = latent_h // patch, latent_w // patch
    img_h = np.linspace(0, gh, gh, endpoint=False, dtype=np.float32)
    img_w = np.linspace(0, gw, gw, endpoint=False, dtype=np.float32)
    img_t = np.linspace(0, num_chunks * frames_per_chunk, num_chunks * frames_per_chunk, endpoint=False, dtype=np.float32)
    cond_h = np.linspace(0, gh, h_queries, endpoint=False, dtype=np.float32)
    cond_w = np.linspace(0, gw, w_queries, endpoint=False, dtype=np.float32)
    cond_t = np.concatenate([
        np.linspace(start_frame_idx + i * frames_per_chunk, start_frame_idx + (i + 1) * frames_per_chunk,
                    vip_frames_per_chunk, endpoint=False, dtype=np.float32)
        for i in range(num_chunks + 1)])
    return (img_t, img_h, img_w), (cond_t, cond_h, cond_w)
