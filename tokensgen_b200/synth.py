"""Synthetic CogVideoX-5b-shaped models and inputs for benchmarks / smoke tests (no checkpoints offline).

Weights: N(0, 0.02) matrices, LayerNorm weight 1, zero-mean small biases — SURVEY.md §8(d) recommends std-0.02 so that
activations stay finite through 42 layers.  Everything is created directly on the GPU in bf16 (a CPU fp32 init of 7 B
parameters would take minutes and 28 GB of host memory).
"""
from __future__ import annotations

import torch

from .transformer import CogVideoXTransformer3DModel

COGVIDEOX_5B = dict(num_attention_heads=48, attention_head_dim=64, in_channels=16, out_channels=16, time_embed_dim=512,
                    text_embed_dim=4096, num_layers=42, patch_size=2, use_rotary_positional_embeddings=True,
                    attention_bias=True)
VIP_5B = dict(length=480, func_type="1", scale=[0.6],
              resampler_params=dict(output_dim=3072, num_height_queries=8, num_width_queries=12, num_temporal_queries=4))


def build_random_model(device="cuda", seed: int = 0, use_vip: bool = True, vip_kwargs=None, **overrides) -> CogVideoXTransformer3DModel:
    cfg = dict(COGVIDEOX_5B)
    cfg.update(overrides)
    with torch.device("meta"):
        m = CogVideoXTransformer3DModel(**cfg)
        if use_vip:
            m.set_vip_layers(None, **(vip_kwargs or VIP_5B))
    m = m.to_empty(device=device).to(torch.bfloat16)
    g = torch.Generator(device=device).manual_seed(seed)
    with torch.no_grad():
        for name, p in m.named_parameters():
            if p.dim() >= 2:
                p.normal_(0.0, 0.02, generator=g)
            elif name.endswith("weight"):
                p.fill_(1.0)
            else:
                p.normal_(0.0, 0.02, generator=g)
    return m.eval()


def window_inputs(B: int = 2, frames: int = 13, channels: int = 16, height: int = 60, width: int = 90, n_text: int = 226,
                  text_dim: int = 4096, vip_frames: int = 5, vip_dim: int = 3072, hq: int = 8, wq: int = 12, seed: int = 42,
                  pin: bool = True):
    """Host-side (pinned) synthetic inputs of one FIFO window: what the reference controller puts on the worker queue."""
    g = torch.Generator().manual_seed(seed)
    mk = lambda *s: torch.randn(*s, generator=g).bfloat16()
    out = dict(latents=mk(1, frames, channels, height, width), old_x0=mk(frames, channels, height, width),
               prompt_embeds=mk(B, n_text, text_dim), image_embeddings=mk(B, vip_frames, vip_dim, hq, wq),
               noise1=mk(1, frames, channels, height, width), noise2=mk(1, frames, channels, height, width))
    if pin and torch.cuda.is_available():
        out = {k: v.pin_memory() for k, v in out.items()}
    return out
