"""ORACLE (test infrastructure, never shipped, never on the product path).

CPU restatement of the reference's video-IP-adapter Resampler (Perceiver resampler: 13 x 1350 patch tokens -> 4 x 8 x 12
condensed tokens), longvgen/video_ipadapter/resampler.py:66-245, as plain functions over a state dict with the
reference's key layout (`latents`, `proj_in.*`, `proj_out.*`, `norm_out.*`, `layers.N.0.{norm1,norm2,to_q,to_kv,to_out,
norm_q,norm_k}.*`, `layers.N.1.net.{0.proj,2}.*`).

Parity pin: tests/test_resampler_cpu.py checks this against tests/golden/resampler_tiny.pt, produced by
oracle/make_goldens.py from the UNMODIFIED reference module.  diffusers' FeedForward is restated (GELU-tanh MLP).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

from oracle.dit import apply_rope

Tensor = torch.Tensor
Rope = Optional[Tuple[Tensor, Tensor]]


@dataclass
class ResamplerConfig:
    """Constructor arguments of Resampler (resampler.py:134-158); defaults = config/infer/edit.yaml:45-58."""
    dim: int = 3072
    depth: int = 4
    dim_head: int = 64
    heads: int = 16
    num_height_queries: int = 8
    num_width_queries: int = 12
    num_temporal_queries: int = 4
    embedding_dim: int = 3072
    output_dim: int = 3072
    ff_mult: int = 4
    max_height_seq_len: int = 30
    max_width_seq_len: int = 45
    max_temporal_seq_len: int = 13

    @property
    def num_latents(self) -> int:
        return self.num_height_queries * self.num_width_queries * self.num_temporal_queries


def resampler_shapes(cfg: ResamplerConfig) -> Dict[str, List[int]]:
    inner = cfg.dim_head * cfg.heads
    s = {"latents": [1, cfg.num_latents, cfg.dim],
         "proj_in.weight": [cfg.dim, cfg.embedding_dim], "proj_in.bias": [cfg.dim],
         "proj_out.weight": [cfg.output_dim, cfg.dim], "proj_out.bias": [cfg.output_dim],
         "norm_out.weight": [cfg.output_dim], "norm_out.bias": [cfg.output_dim]}
    for i in range(cfg.depth):
        a, f = f"layers.{i}.0", f"layers.{i}.1"
        for n in ("norm1", "norm2"):
            s[f"{a}.{n}.weight"], s[f"{a}.{n}.bias"] = [cfg.dim], [cfg.dim]
        s[f"{a}.to_q.weight"] = [inner, cfg.dim]
        s[f"{a}.to_kv.weight"] = [2 * inner, cfg.dim]
        s[f"{a}.to_out.weight"] = [cfg.dim, inner]
        for n in ("norm_q", "norm_k"):
            s[f"{a}.{n}.weight"], s[f"{a}.{n}.bias"] = [cfg.dim_head], [cfg.dim_head]
        s[f"{f}.net.0.proj.weight"], s[f"{f}.net.0.proj.bias"] = [4 * cfg.dim, cfg.dim], [4 * cfg.dim]
        s[f"{f}.net.2.weight"], s[f"{f}.net.2.bias"] = [cfg.dim, 4 * cfg.dim], [cfg.dim]
    return s


def _lin(x: Tensor, sd, key: str, dtype) -> Tensor:
    b = sd.get(key + ".bias")
    return F.linear(x, sd[key + ".weight"].to(dtype), None if b is None else b.to(dtype))


def _ln(x: Tensor, sd, key: str, eps: float, dtype) -> Tensor:
    return F.layer_norm(x, (x.shape[-1],), sd[key + ".weight"].to(dtype), sd[key + ".bias"].to(dtype), eps)


def perceiver_attention(sd, pre: str, cfg: ResamplerConfig, x: Tensor, latents: Tensor, image_rope: Rope,
                        sampling_rope: Rope, dtype) -> Tensor:
    """PerceiverAttention.forward, resampler.py:86-129.  K/V = [x; latents]; RoPE on the x keys with the 13x30x45 grid and
    on the queries / latent keys with the sampled grid that starts at t = 1000."""
    x = _ln(x, sd, pre + ".norm1", 1e-5, dtype)
    latents = _ln(latents, sd, pre + ".norm2", 1e-5, dtype)
    b, l, _ = latents.shape
    q = _lin(latents, sd, pre + ".to_q", dtype)
    k, v = _lin(torch.cat((x, latents), dim=-2), sd, pre + ".to_kv", dtype).chunk(2, dim=-1)
    heads = lambda t: t.view(b, t.shape[1], cfg.heads, -1).transpose(1, 2)
    q, k, v = heads(q), heads(k), heads(v)
    q = _ln(q, sd, pre + ".norm_q", 1e-6, dtype)
    k = _ln(k, sd, pre + ".norm_k", 1e-6, dtype)
    if image_rope is not None:
        k = torch.cat([apply_rope(k[:, :, :-l], image_rope), k[:, :, -l:]], dim=2)
    if sampling_rope is not None:
        q = apply_rope(q, sampling_rope)
        k = torch.cat([k[:, :, :-l], apply_rope(k[:, :, -l:], sampling_rope)], dim=2)
    out = F.scaled_dot_product_attention(q, k, v, scale=1 / math.sqrt(cfg.dim_head))
    out = out.permute(0, 2, 1, 3).reshape(b, l, -1)
    return _lin(out, sd, pre + ".to_out", dtype)


def feed_forward(sd, pre: str, x: Tensor, dtype) -> Tensor:
    """diffusers FeedForward("gelu-approximate"): Linear -> GELU(tanh) -> Linear (dropout p = 0)."""
    h = F.gelu(_lin(x, sd, pre + ".net.0.proj", dtype), approximate="tanh")
    return _lin(h, sd, pre + ".net.2", dtype)


def resampler_forward(sd: Dict[str, Tensor], cfg: ResamplerConfig, x: Tensor, image_rope: Rope, sampling_rope: Rope,
                      dtype=torch.float32, pca=None) -> Tensor:
    """Resampler.forward, resampler.py:209-244.  x: [b, f, n, embedding_dim] -> [b, num_temporal_queries, output_dim, h, w]."""
    b, f, n, _ = x.shape
    x = _lin(x.to(dtype).reshape(b * f, n, -1), sd, "proj_in", dtype).reshape(b, f * n, -1)
    latents = sd["latents"].to(dtype).expand(b, -1, -1)
    for i in range(cfg.depth):
        latents = perceiver_attention(sd, f"layers.{i}.0", cfg, x, latents, image_rope, sampling_rope, dtype) + latents
        latents = feed_forward(sd, f"layers.{i}.1", latents, dtype) + latents
    latents = _ln(_lin(latents, sd, "proj_out", dtype), sd, "norm_out", 1e-5, dtype)
    if pca is not None:  # resampler.py:230-237: keep the first 16 principal components
        flat = latents.reshape(-1, latents.shape[-1]).to(pca.components_.dtype)
        t = pca.transform(flat)
        t[:, 16:] = 0.0
        latents = pca.inverse_transform(t).reshape(b, -1, latents.shape[-1]).to(dtype)
    return latents.reshape(b, cfg.num_temporal_queries, cfg.num_height_queries, cfg.num_width_queries, -1).permute(0, 1, 4, 2, 3)
