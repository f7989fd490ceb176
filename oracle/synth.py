"""ORACLE (test infrastructure): deterministic synthetic weights / inputs shared by the golden generator, the tests,
smoke() and bench.py, so that golden files only need to carry shapes, a checksum, inputs and reference outputs."""
from __future__ import annotations

import hashlib
from typing import Dict, List

import torch


def synth_state_dict(shapes: Dict[str, List[int]], seed: int, dtype=torch.bfloat16) -> Dict[str, torch.Tensor]:
    """Values depend only on (sorted key order, shapes, seed): LayerNorm weights ~ 1 + 0.1 N(0,1), biases ~ 0.05 N(0,1),
    matrices ~ N(0,1)/sqrt(fan_in).  Returned in `dtype` (bf16 = what the product path stores)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k in sorted(shapes):
        shp = list(shapes[k])
        x = torch.randn(shp, generator=g, dtype=torch.float32)
        if len(shp) == 1:
            is_norm_w = k.endswith(".weight") and (".norm" in k or k.startswith("norm"))
            x = 1.0 + 0.1 * x if is_norm_w else 0.05 * x
        else:
            fan_in = 1
            for s in shp[1:]:
                fan_in *= s
            x = x / fan_in ** 0.5
        sd[k] = x.to(dtype)
    return sd


def state_dict_digest(sd: Dict[str, torch.Tensor]) -> str:
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(sd[k].detach().cpu().contiguous().view(torch.uint8).numpy().tobytes())
    return h.hexdigest()


def dit_shapes(heads: int, head_dim: int, layers: int, time_dim: int, text_dim: int, in_ch: int, out_ch: int, patch: int,
               vip_dim: int, use_vip: bool = True) -> Dict[str, List[int]]:
    """State-dict key layout of CogVideoXTransformer3DModel + set_vip_layers(func_type "1") (verified
    against the instantiated reference by tests/golden/dit_tiny.json's key list)."""
    d = heads * head_dim
    s: Dict[str, List[int]] = {
        "patch_embed.proj.weight": [d, in_ch, patch, patch], "patch_embed.proj.bias": [d],
        "patch_embed.text_proj.weight": [d, text_dim], "patch_embed.text_proj.bias": [d],
        "time_embedding.linear_1.weight": [time_dim, d], "time_embedding.linear_1.bias": [time_dim],
        "time_embedding.linear_2.weight": [time_dim, time_dim], "time_embedding.linear_2.bias": [time_dim],
        "norm_final.weight": [d], "norm_final.bias": [d],
        "norm_out.linear.weight": [2 * d, time_dim], "norm_out.linear.bias": [2 * d],
        "norm_out.norm.weight": [d], "norm_out.norm.bias": [d],
        "proj_out.weight": [patch * patch * out_ch, d], "proj_out.bias": [patch * patch * out_ch],
    }
    if use_vip:
        s["patch_embed.vip_proj.weight"] = [d, vip_dim]
        s["patch_embed.vip_proj.bias"] = [d]
    for i in range(layers):
        p = f"transformer_blocks.{i}."
        for n, mult in (("norm1", 6), ("norm2", 6)) + ((("vip_norm1", 3), ("vip_norm2", 3)) if use_vip else ()):
            s[p + n + ".linear.weight"] = [mult * d, time_dim]
            s[p + n + ".linear.bias"] = [mult * d]
            s[p + n + ".norm.weight"] = [d]
            s[p + n + ".norm.bias"] = [d]
        for n in ("to_q", "to_k", "to_v", "to_out.0"):
            s[p + "attn1." + n + ".weight"] = [d, d]
            s[p + "attn1." + n + ".bias"] = [d]
        for n in ("norm_q", "norm_k"):
            s[p + "attn1." + n + ".weight"] = [head_dim]
            s[p + "attn1." + n + ".bias"] = [head_dim]
        if use_vip:
            for n in ("vip_to_q", "vip_to_k", "vip_to_v"):
                s[p + "attn1.processor." + n + ".weight"] = [d, d]
                s[p + "attn1.processor." + n + ".bias"] = [d]
            for n in ("vip_norm_q", "vip_norm_k"):
                s[p + "attn1.processor." + n + ".weight"] = [head_dim]
                s[p + "attn1.processor." + n + ".bias"] = [head_dim]
        s[p + "ff.net.0.proj.weight"] = [4 * d, d]
        s[p + "ff.net.0.proj.bias"] = [4 * d]
        s[p + "ff.net.2.weight"] = [d, 4 * d]
        s[p + "ff.net.2.bias"] = [d]
    return s


def conditioning_clip(frames: int = 10, height: int = 480, width: int = 720) -> torch.Tensor:
    """Deterministic [1, F, 3, H, W] conditioning video in [-1, 1] (moving gradients + a seeded low-amplitude texture),
    generated instead of stored: 10 frames of 480 x 720 fp32 would be a 41 MB fixture."""
    t = torch.arange(frames, dtype=torch.float32)[:, None, None]
    y = torch.arange(height, dtype=torch.float32)[None, :, None]
    x = torch.arange(width, dtype=torch.float32)[None, None, :]
    r = torch.sin(0.031 * x + 0.4 * t) * torch.cos(0.017 * y)
    g = torch.sin(0.023 * y - 0.3 * t + 1.0)
    b = torch.cos(0.011 * (x + y) + 0.2 * t)
    clip = torch.stack(torch.broadcast_tensors(r, g, b), dim=1)
    noise = torch.rand(clip.shape, generator=torch.Generator().manual_seed(5)) * 0.2 - 0.1
    return (clip * 0.9 + noise).clamp(-1, 1).unsqueeze(0)


def keyed_noise(key, shape, dtype=torch.bfloat16) -> torch.Tensor:
    """Deterministic N(0,1) draw identified by a tuple of small non-negative ints (iteration, window start, frame, draw...):
    the FIFO-stage golden replaces the reference's global-RNG draws with these so that the product path can be fed the very
    same noise on any device (tests/golden/fifo_stage_tiny.pt)."""
    seed = 0x5EED
    for k in key:
        seed = (seed * 1000003 + int(k) + 1) % (1 << 62)
    return torch.randn(tuple(shape), generator=torch.Generator().manual_seed(seed), dtype=torch.float32).to(dtype)
