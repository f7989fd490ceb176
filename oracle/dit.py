"""ORACLE (test infrastructure, never shipped, never on the product path).

CPU restatement of the reference's CogVideoX-5b DiT window forward with the video-IP-adapter (func_type "1"), written
as plain functions over a state dict with the reference's key layout.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this module.

Parity pin: tests/test_oracle_cpu.py checks every function here against tests/golden/*.pt, which
oracle/make_goldens.py produced by running the UNMODIFIED reference modules (/root/reference, imported through
oracle/stubs) on seeded inputs in this container.  The arithmetic of diffusers' FeedForward is restated from its
published semantics (diffusers 0.31.0.dev0 is a pip dependency of the reference, environment.yml:58, absent offline).

Each function cites the reference lines it follows.  `dtype` selects the compute type: torch.float32 (the accuracy
reference for the bf16 CUDA kernels) or torch.bfloat16 (what the reference itself runs; used for the CPU baseline).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
Rope = Optional[Tuple[Tensor, Tensor]]


@dataclass
class DitConfig:
    """The subset of CogVideoXTransformer3DModel's config the forward depends on (cogvideox_transformer_3d.py:392-420)."""
    num_attention_heads: int = 48
    attention_head_dim: int = 64
    in_channels: int = 16
    out_channels: int = 16
    time_embed_dim: int = 512
    text_embed_dim: int = 4096
    num_layers: int = 42
    patch_size: int = 2
    flip_sin_to_cos: bool = True
    freq_shift: float = 0.0
    norm_eps: float = 1e-5
    qk_eps: float = 1e-6
    vip_length: int = 480       # (num_temporal_queries + 1) * num_height_queries * num_width_queries
    vip_embed_dim: int = 3072   # resampler output_dim
    vip_scale: float = 0.6
    use_vip: bool = True

    @property
    def inner_dim(self) -> int:
        return self.num_attention_heads * self.attention_head_dim


# ----------------------------------------------------------------------------------------------- small pieces
def timestep_sinusoid(timesteps: Tensor, dim: int, flip_sin_to_cos: bool, freq_shift: float) -> Tensor:
    """get_timestep_embedding, embeddings.py:28-79 (max_period 10000, scale 1)."""
    half = dim // 2
    exponent = -math.log(10000) * torch.arange(0, half, dtype=torch.float32) / (half - freq_shift)
    arg = timesteps[:, None].float() * torch.exp(exponent)[None, :]
    emb = torch.cat([torch.sin(arg), torch.cos(arg)], dim=-1)
    if flip_sin_to_cos:
        emb = torch.cat([emb[:, half:], emb[:, :half]], dim=-1)
    return emb


def linear(x: Tensor, sd: Dict[str, Tensor], key: str, dtype) -> Tensor:
    b = sd.get(key + ".bias")
    return F.linear(x, sd[key + ".weight"].to(dtype), None if b is None else b.to(dtype))


def time_embedding(sd, cfg: DitConfig, timestep: Tensor, batch: int, dtype) -> Tensor:
    """cogvideox_transformer_3d.py:669-680: sinusoid -> cast -> linear_1 -> SiLU -> linear_2, reshaped [B, F, C]."""
    t = timestep.reshape(-1)
    t_emb = timestep_sinusoid(t, cfg.inner_dim, cfg.flip_sin_to_cos, cfg.freq_shift).to(dtype)
    h = F.silu(linear(t_emb, sd, "time_embedding.linear_1", dtype))
    emb = linear(h, sd, "time_embedding.linear_2", dtype)
    return emb.reshape(batch, -1, emb.shape[-1])


def patch_embed(sd, cfg: DitConfig, text: Tensor, latents: Tensor, vip: Optional[Tensor], dtype) -> Tensor:
    """CogVideoXPatchEmbed.forward, embeddings.py:502-544 (RoPE model: no additive positional embedding).
    Returns [B, n_text + F*h*w + n_vip, d] in the order [text, video, vip]."""
    text_e = linear(text.to(dtype), sd, "patch_embed.text_proj", dtype)
    B, Fr, C, H, W = latents.shape
    p = cfg.patch_size
    x = latents.to(dtype).reshape(B * Fr, C, H, W)
    x = F.conv2d(x, sd["patch_embed.proj.weight"].to(dtype), sd["patch_embed.proj.bias"].to(dtype), stride=p)
    x = x.reshape(B, Fr, x.shape[1], -1).transpose(2, 3).reshape(B, -1, x.shape[1])
    parts = [text_e, x]
    if vip is not None:
        v = vip.to(dtype).permute(0, 1, 3, 4, 2).reshape(B, -1, vip.shape[2])  # "b f c h w -> b (f h w) c"
        parts.append(linear(v, sd, "patch_embed.vip_proj", dtype))
    return torch.cat(parts, dim=1)


def apply_rope(x: Tensor, rope: Tuple[Tensor, Tensor]) -> Tensor:
    """apply_rotary_emb (use_real, unbind_dim=-1), embeddings.py:868-884: fp32 math on interleaved pairs, cast back."""
    cos, sin = rope
    cos, sin = cos[None, None].float(), sin[None, None].float()
    x0, x1 = x.reshape(*x.shape[:-1], -1, 2).unbind(-1)
    rot = torch.stack([-x1, x0], dim=-1).flatten(3)
    return (x.float() * cos + rot.float() * sin).to(x.dtype)


def layer_norm(x: Tensor, sd, key: str, eps: float, dtype) -> Tensor:
    return F.layer_norm(x, (x.shape[-1],), sd[key + ".weight"].to(dtype), sd[key + ".bias"].to(dtype), eps)


def layernorm_zero(sd, prefix: str, hidden: Tensor, enc: Tensor, temb: Tensor, eps: float, dtype):
    """CogVideoXLayerNormZero.forward, normalization.py:443-460: video rows use their frame's vectors, text rows frame 0."""
    B, Fr, _ = temb.shape
    hw = hidden.shape[1] // Fr
    mod = linear(F.silu(temb.reshape(B * Fr, -1)), sd, prefix + ".linear", dtype)
    shift, scale, gate, e_shift, e_scale, e_gate = [c.reshape(B, Fr, -1) for c in mod.chunk(6, dim=1)]
    rep = lambda t: t.repeat_interleave(hw, dim=1)  # "b f c -> b (f hw) c"
    h = layer_norm(hidden, sd, prefix + ".norm", eps, dtype) * (1 + rep(scale)) + rep(shift)
    e = layer_norm(enc, sd, prefix + ".norm", eps, dtype) * (1 + e_scale)[:, [0]] + e_shift[:, [0]]
    return h, e, rep(gate), e_gate[:, [0]]


def vip_layernorm_zero(sd, prefix: str, vip: Tensor, temb: Tensor, eps: float, dtype):
    """CogVideoXVIPLayerNormZero.forward, normalization.py:477-488."""
    B, Fr, _ = temb.shape
    mod = linear(F.silu(temb.reshape(B * Fr, -1)), sd, prefix + ".linear", dtype)
    e_shift, e_scale, e_gate = [c.reshape(B, Fr, -1) for c in mod.chunk(3, dim=1)]
    v = layer_norm(vip, sd, prefix + ".norm", eps, dtype) * (1 + e_scale)[:, [0]] + e_shift[:, [0]]
    return v, e_gate[:, [0]]


def sdpa(q: Tensor, k: Tensor, v: Tensor) -> Tensor:
    """F.scaled_dot_product_attention, non-causal, scale 1/sqrt(head_dim) (attention_processor.py:2067-2069)."""
    return F.scaled_dot_product_attention(q, k, v, dropout_p=0.0, is_causal=False)


def feed_forward(sd, prefix: str, x: Tensor, dtype) -> Tensor:
    """diffusers FeedForward("gelu-approximate"): Linear -> GELU(tanh) -> Linear (dropout p=0)."""
    h = F.gelu(linear(x, sd, prefix + ".net.0.proj", dtype), approximate="tanh")
    return linear(h, sd, prefix + ".net.2", dtype)


# ----------------------------------------------------------------------------------------------- attention
def vip_attention(sd, pre: str, cfg: DitConfig, hidden: Tensor, enc: Tensor, rope: Rope, vip_img_rope: Rope,
                  vip_cond_rope: Rope, dtype):
    """VideoIPAdapterCogVideoXAttnProcessor2_0.__call__, attention_processor.py:1982-2155.
    hidden [B, n_video, d]; enc [B, n_text + n_vip, d].  Returns (hidden_out, enc_out) in the same split."""
    H, D = cfg.num_attention_heads, cfg.attention_head_dim
    n_vip = cfg.vip_length
    text, vip = enc[:, :-n_vip], enc[:, -n_vip:]
    n_text = text.shape[1]
    tv = torch.cat([text, hidden], dim=1)
    B = tv.shape[0]
    heads = lambda t: t.reshape(B, -1, H, D).transpose(1, 2)
    proc = pre + ".processor"

    q = heads(linear(tv, sd, pre + ".to_q", dtype))
    k = heads(linear(tv, sd, pre + ".to_k", dtype))
    v = heads(linear(tv, sd, pre + ".to_v", dtype))
    q_tv = heads(linear(tv, sd, proc + ".vip_to_q", dtype))
    k_tv = heads(linear(tv, sd, proc + ".vip_to_k", dtype))
    v_tv = heads(linear(tv, sd, proc + ".vip_to_v", dtype))
    q_vip = heads(linear(vip, sd, proc + ".vip_to_q", dtype))
    k_vip = heads(linear(vip, sd, proc + ".vip_to_k", dtype))
    v_vip = heads(linear(vip, sd, proc + ".vip_to_v", dtype))

    q = layer_norm(q, sd, pre + ".norm_q", cfg.qk_eps, dtype)
    k = layer_norm(k, sd, pre + ".norm_k", cfg.qk_eps, dtype)
    q_tv = layer_norm(q_tv, sd, proc + ".vip_norm_q", cfg.qk_eps, dtype)
    q_vip = layer_norm(q_vip, sd, proc + ".vip_norm_q", cfg.qk_eps, dtype)
    k_tv = layer_norm(k_tv, sd, proc + ".vip_norm_k", cfg.qk_eps, dtype)
    k_vip = layer_norm(k_vip, sd, proc + ".vip_norm_k", cfg.qk_eps, dtype)

    if rope is not None:  # :2043-2056 — text rows are never rotated
        q = torch.cat([q[:, :, :n_text], apply_rope(q[:, :, n_text:], rope)], dim=2)
        k = torch.cat([k[:, :, :n_text], apply_rope(k[:, :, n_text:], rope)], dim=2)
        q_tv = torch.cat([q_tv[:, :, :n_text], apply_rope(q_tv[:, :, n_text:], vip_img_rope)], dim=2)
        k_tv = torch.cat([k_tv[:, :, :n_text], apply_rope(k_tv[:, :, n_text:], vip_img_rope)], dim=2)
        q_vip = apply_rope(q_vip, vip_cond_rope)
        k_vip = apply_rope(k_vip, vip_cond_rope)

    out = sdpa(q, k, v)                                                      # :2067
    cross = sdpa(q_tv, k_vip, v_vip)                                         # :2117
    vip_out = sdpa(q_vip, torch.cat([k_tv, k_vip], 2), torch.cat([v_tv, v_vip], 2))  # :2120
    scale = torch.tensor(cfg.vip_scale).to(device=out.device, dtype=out.dtype)  # :2126-2133 (list of one -> scalar)
    out = out + scale * cross
    out = torch.cat([out, vip_out], dim=2).transpose(1, 2).reshape(B, -1, H * D)
    out = linear(out, sd, pre + ".to_out.0", dtype)
    text_o, hid_o, vip_o = out.split([n_text, out.shape[1] - n_text - n_vip, n_vip], dim=1)
    return hid_o, torch.cat([text_o, vip_o], dim=1)


def plain_attention(sd, pre: str, cfg: DitConfig, hidden: Tensor, enc: Tensor, rope: Rope, dtype):
    """CogVideoXAttnProcessor2_0.__call__, attention_processor.py:1895-1953."""
    H, D = cfg.num_attention_heads, cfg.attention_head_dim
    n_text = enc.shape[1]
    tv = torch.cat([enc, hidden], dim=1)
    B = tv.shape[0]
    heads = lambda t: t.reshape(B, -1, H, D).transpose(1, 2)
    q = layer_norm(heads(linear(tv, sd, pre + ".to_q", dtype)), sd, pre + ".norm_q", cfg.qk_eps, dtype)
    k = layer_norm(heads(linear(tv, sd, pre + ".to_k", dtype)), sd, pre + ".norm_k", cfg.qk_eps, dtype)
    v = heads(linear(tv, sd, pre + ".to_v", dtype))
    if rope is not None:
        q = torch.cat([q[:, :, :n_text], apply_rope(q[:, :, n_text:], rope)], dim=2)
        k = torch.cat([k[:, :, :n_text], apply_rope(k[:, :, n_text:], rope)], dim=2)
    out = sdpa(q, k, v).transpose(1, 2).reshape(B, -1, H * D)
    out = linear(out, sd, pre + ".to_out.0", dtype)
    return out[:, n_text:], out[:, :n_text]


# ----------------------------------------------------------------------------------------------- block / model
def block_forward(sd, pre: str, cfg: DitConfig, hidden: Tensor, enc: Tensor, temb: Tensor, rope: Rope,
                  vip_img_rope: Rope, vip_cond_rope: Rope, dtype):
    """CogVideoXBlock.forward, cogvideox_transformer_3d.py:221-332 (func_type "1" when cfg.use_vip)."""
    eps = cfg.norm_eps
    if cfg.use_vip:
        text, vip = enc[:, :-cfg.vip_length], enc[:, -cfg.vip_length:]
    else:
        text, vip = enc, None
    n_text = text.shape[1]

    nh, ne, gate, e_gate = layernorm_zero(sd, pre + ".norm1", hidden, text, temb, eps, dtype)
    if cfg.use_vip:
        nv, v_gate = vip_layernorm_zero(sd, pre + ".vip_norm1", vip, temb, eps, dtype)
        a_h, a_e = vip_attention(sd, pre + ".attn1", cfg, nh, torch.cat([ne, nv], 1), rope, vip_img_rope,
                                 vip_cond_rope, dtype)
        a_t, a_v = a_e[:, :n_text], a_e[:, n_text:]
    else:
        a_h, a_t = plain_attention(sd, pre + ".attn1", cfg, nh, ne, rope, dtype)
    hidden = hidden + gate * a_h
    text = text + e_gate * a_t
    if cfg.use_vip:
        vip = vip + v_gate * a_v

    nh, ne, gate_ff, e_gate_ff = layernorm_zero(sd, pre + ".norm2", hidden, text, temb, eps, dtype)
    ff = feed_forward(sd, pre + ".ff", torch.cat([ne, nh], dim=1), dtype)
    hidden = hidden + gate_ff * ff[:, n_text:]
    text = text + e_gate_ff * ff[:, :n_text]
    if cfg.use_vip:
        nv, v_gate_ff = vip_layernorm_zero(sd, pre + ".vip_norm2", vip, temb, eps, dtype)
        vip = vip + v_gate_ff * feed_forward(sd, pre + ".ff", nv, dtype)
        return hidden, torch.cat([text, vip], dim=1)
    return hidden, text


def final_layers(sd, cfg: DitConfig, hidden: Tensor, temb: Tensor, latent_shape, dtype) -> Tensor:
    """norm_final -> AdaLayerNorm (shift, scale order) -> proj_out -> unpatchify.
    cogvideox_transformer_3d.py:736-759; normalization.py:70-92.  `hidden` = video rows only (LayerNorm is per row,
    so dropping the text/vip rows before or after norm_final is the same)."""
    B, Fr, C, H, W = latent_shape
    p = cfg.patch_size
    Ft = temb.shape[1]  # 1 when the timestep is per sample, F when it is per frame
    hw = hidden.shape[1] // Ft
    x = layer_norm(hidden, sd, "norm_final", cfg.norm_eps, dtype)
    mod = linear(F.silu(temb.reshape(B * Ft, -1)), sd, "norm_out.linear", dtype)
    shift, scale = [c.reshape(B, Ft, -1).repeat_interleave(hw, dim=1) for c in mod.chunk(2, dim=1)]
    x = layer_norm(x, sd, "norm_out.norm", cfg.norm_eps, dtype) * (1 + scale) + shift
    x = linear(x, sd, "proj_out", dtype)
    x = x.reshape(B, Fr, H // p, W // p, -1, p, p)
    return x.permute(0, 1, 4, 2, 5, 3, 6).flatten(5, 6).flatten(3, 4)


def dit_forward(sd: Dict[str, Tensor], cfg: DitConfig, latents: Tensor, text: Tensor, timestep: Tensor,
                vip: Optional[Tensor], rope: Rope, vip_img_rope: Rope = None, vip_cond_rope: Rope = None,
                dtype=torch.float32) -> Tensor:
    """CogVideoXTransformer3DModel.forward, cogvideox_transformer_3d.py:636-770.
    latents [B,F,C,H,W]; text [B,n_text,text_dim]; timestep [B] or [B,F]; vip [B,f_vip,vip_dim,h_q,w_q]."""
    B, Fr = latents.shape[:2]
    if timestep.dim() == 1:
        timestep = timestep[:, None]
    temb = time_embedding(sd, cfg, timestep, B, dtype)           # [B, F or 1, C]
    x = patch_embed(sd, cfg, text, latents, vip if cfg.use_vip else None, dtype)
    n_text = text.shape[1]
    if cfg.use_vip:
        enc = torch.cat([x[:, :n_text], x[:, -cfg.vip_length:]], dim=1)
        hidden = x[:, n_text:-cfg.vip_length]
    else:
        enc, hidden = x[:, :n_text], x[:, n_text:]
    for i in range(cfg.num_layers):
        hidden, enc = block_forward(sd, f"transformer_blocks.{i}", cfg, hidden, enc, temb, rope, vip_img_rope,
                                    vip_cond_rope, dtype)
    return final_layers(sd, cfg, hidden, temb, latents.shape, dtype)
