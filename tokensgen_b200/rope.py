"""3D-RoPE tables and position grids (host side), mirroring longvgen/models/embeddings.py:571-707,774-837 and the grid
builders of longvgen/pipeline/pipeline_cogvideox_mp_fifo.py:769-813,1061-1103.

The per-axis 1-D tables (13x16, 30x24, 45x24 entries) are computed on the host with the same torch fp32 ops as the
reference — so they are bit-identical — and only expanded to the [tokens, 64] layout on the device (pure data
movement).  The reference instead builds the full 17 550 x 64 tables in numpy/torch on the CPU and uploads 9 MB per
window (cogvideo_sampling_mp_fifo.py:478-489).
"""
from __future__ import annotations

from typing import Tuple

import numpy as np
import torch


def get_1d_rotary_pos_embed(dim: int, pos, theta: float = 10000.0) -> Tuple[torch.Tensor, torch.Tensor]:
    """use_real=True, repeat_interleave_real=True branch (embeddings.py:812-825)."""
    assert dim % 2 == 0
    if isinstance(pos, int):
        pos = torch.arange(pos)
    if isinstance(pos, np.ndarray):
        pos = torch.from_numpy(pos)
    freqs = 1.0 / (theta ** (torch.arange(0, dim, 2, dtype=torch.float32)[: (dim // 2)] / dim))
    freqs = torch.outer(pos, freqs)
    return freqs.cos().repeat_interleave(2, dim=1).float(), freqs.sin().repeat_interleave(2, dim=1).float()


def get_3d_rotary_pos_embed_v2(embed_dim, grid_t, grid_h, grid_w, dim_t=None, dim_h=None, dim_w=None, device="cpu"):
    """[T*H*W, embed_dim] cos and sin tables for explicit per-axis position grids (embeddings.py:641-707)."""
    dim_t = embed_dim // 4 if dim_t is None else dim_t
    dim_h = embed_dim // 8 * 3 if dim_h is None else dim_h
    dim_w = embed_dim // 8 * 3 if dim_w is None else dim_w
    T, H, W = len(grid_t), len(grid_h), len(grid_w)
    axes = [get_1d_rotary_pos_embed(dim_t, np.asarray(grid_t)), get_1d_rotary_pos_embed(dim_h, np.asarray(grid_h)),
            get_1d_rotary_pos_embed(dim_w, np.asarray(grid_w))]
    out = []
    for which in (0, 1):
        t = axes[0][which].to(device)[:, None, None, :].expand(T, H, W, dim_t)
        h = axes[1][which].to(device)[None, :, None, :].expand(T, H, W, dim_h)
        w = axes[2][which].to(device)[None, None, :, :].expand(T, H, W, dim_w)
        out.append(torch.cat([t, h, w], dim=-1).reshape(T * H * W, -1).contiguous())
    return out[0], out[1]


def get_3d_rotary_pos_embed(embed_dim, crops_coords, grid_size, device="cpu"):
    """embeddings.py:571-639: fp32 linspace grids between the crop corners."""
    start, stop = crops_coords
    T, H, W = grid_size
    gt = np.linspace(start[0], stop[0], T, endpoint=False, dtype=np.float32)
    gh = np.linspace(start[1], stop[1], H, endpoint=False, dtype=np.float32)
    gw = np.linspace(start[2], stop[2], W, endpoint=False, dtype=np.float32)
    return get_3d_rotary_pos_embed_v2(embed_dim, gt, gh, gw, device=device)


def vip_position_grids(latent_h: int, latent_w: int, patch: int, num_chunks: int, frames_per_chunk: int,
                       vip_frames_per_chunk: int, h_queries: int, w_queries: int, start_frame_idx: float):
    """pipeline_cogvideox_mp_fifo.py:1061-1103 -> ((img_t, img_h, img_w), (cond_t, cond_h, cond_w)), all fp32 numpy."""
    gh, gw = latent_h // patch, latent_w // patch
    img = (np.linspace(0, num_chunks * frames_per_chunk, num_chunks * frames_per_chunk, endpoint=False, dtype=np.float32),
           np.linspace(0, gh, gh, endpoint=False, dtype=np.float32), np.linspace(0, gw, gw, endpoint=False, dtype=np.float32))
    cond_t = np.concatenate([
        np.linspace(start_frame_idx + i * frames_per_chunk, start_frame_idx + (i + 1) * frames_per_chunk,
                    vip_frames_per_chunk, endpoint=False, dtype=np.float32) for i in range(num_chunks + 1)])
    cond = (cond_t, np.linspace(0, gh, h_queries, endpoint=False, dtype=np.float32),
            np.linspace(0, gw, w_queries, endpoint=False, dtype=np.float32))
    return img, cond
