"""Host-side mirror of the reference's DiT module / attention-processor API, executing on the C-ABI CUDA library.

Mirrors (same class names, constructor arguments, attribute names, state-dict keys and call signatures):
  longvgen/models/cogvideox_transformer_3d.py : CogVideoXBlock (:54-332), CogVideoXTransformer3DModel (:335-770)
  longvgen/models/attention_processor.py      : Attention (:46-717, the subset this path uses),
                                                CogVideoXAttnProcessor2_0 (:1885-1953),
                                                VideoIPAdapterCogVideoXAttnProcessor2_0 (:1955-2155)
  longvgen/models/normalization.py            : CogVideoXLayerNormZero (:426-460), CogVideoXVIPLayerNormZero (:462-488),
                                                AdaLayerNorm (:34-92)
  longvgen/models/embeddings.py               : CogVideoXPatchEmbed (:380-568), Timesteps / TimestepEmbedding (:920-984)
  diffusers.models.attention.FeedForward      ("gelu-approximate")

nn.Module is used as the parameter container (so `load_state_dict` of a reference checkpoint / vip.pt works key for
key); every tensor op of the forward pass is a tg_* call.  There is no PyTorch fallback: without the native library or
a CUDA device the forward raises.
"""
from __future__ import annotations

import inspect
import os
from types import SimpleNamespace
from typing import Any, Dict, Optional, Tuple, Union

import numpy as np
import torch
import torch.nn as nn

from . import _ext as E


def _bf16_scalar(x: float) -> float:
    """The reference materialises the vip scale as a bf16 tensor (attention_processor.py:2126-2129)."""
    return float(torch.tensor(float(x)).bfloat16())


def _rope_pair(rope, device):
    if rope is None:
        return None
    cos, sin = rope
    cos = cos.to(device=device, dtype=torch.float32).contiguous()
    sin = sin.to(device=device, dtype=torch.float32).contiguous()
    return cos, sin


# =================================================================================================== parameter containers
class Timesteps(nn.Module):
    def __init__(self, num_channels: int, flip_sin_to_cos: bool, downscale_freq_shift: float, scale: int = 1):
        super().__init__()
        self.num_channels, self.flip_sin_to_cos, self.downscale_freq_shift, self.scale = \
            num_channels, flip_sin_to_cos, downscale_freq_shift, scale


class TimestepEmbedding(nn.Module):
    def __init__(self, in_channels: int, time_embed_dim: int, act_fn: str = "silu"):
        super().__init__()
        if act_fn != "silu":
            raise NotImplementedError("timestep_activation_fn must be 'silu' (CogVideoX)")
        self.linear_1 = nn.Linear(in_channels, time_embed_dim)
        self.linear_2 = nn.Linear(time_embed_dim, time_embed_dim)


class CogVideoXPatchEmbed(nn.Module):
    def __init__(self, patch_size=2, in_channels=16, embed_dim=1920, text_embed_dim=4096, bias=True, **_unused):
        super().__init__()
        self.patch_size, self.embed_dim, self.use_vip = patch_size, embed_dim, False
        self.proj = nn.Conv2d(in_channels, embed_dim, kernel_size=(patch_size, patch_size), stride=patch_size, bias=bias)
        self.text_proj = nn.Linear(text_embed_dim, embed_dim)

    def set_vip_layers(self, **kwargs):
        self.use_vip = True
        self.func_type = kwargs.get("func_type", "1")
        rp = kwargs["resampler_params"]
        self.vip_proj = nn.Linear(rp["output_dim"], self.embed_dim)
        self.vip_num_height_queries = rp["num_height_queries"]
        self.vip_num_width_queries = rp["num_width_queries"]
        self.vip_num_temporal_queries = rp["num_temporal_queries"]


class _LayerNormZeroBase(nn.Module):
    def __init__(self, conditioning_dim, embedding_dim, chunks, elementwise_affine=True, eps=1e-5, bias=True):
        super().__init__()
        self.silu = nn.SiLU()
        self.linear = nn.Linear(conditioning_dim, chunks * embedding_dim, bias=bias)
        self.norm = nn.LayerNorm(embedding_dim, eps=eps, elementwise_affine=elementwise_affine)


class CogVideoXLayerNormZero(_LayerNormZeroBase):
    def __init__(self, conditioning_dim, embedding_dim, elementwise_affine=True, eps=1e-5, bias=True):
        super().__init__(conditioning_dim, embedding_dim, 6, elementwise_affine, eps, bias)


class CogVideoXVIPLayerNormZero(_LayerNormZeroBase):
    def __init__(self, conditioning_dim, embedding_dim, elementwise_affine=True, eps=1e-5, bias=True):
        super().__init__(conditioning_dim, embedding_dim, 3, elementwise_affine, eps, bias)


class AdaLayerNorm(nn.Module):
    def __init__(self, embedding_dim, output_dim=None, norm_elementwise_affine=False, norm_eps=1e-5, chunk_dim=0, **_unused):
        super().__init__()
        self.chunk_dim = chunk_dim
        output_dim = output_dim or embedding_dim * 2
        self.silu = nn.SiLU()
        self.linear = nn.Linear(embedding_dim, output_dim)
        self.norm = nn.LayerNorm(output_dim // 2, norm_eps, norm_elementwise_affine)


class _GELUProj(nn.Module):
    def __init__(self, dim_in, dim_out, bias=True):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out, bias=bias)


class FeedForward(nn.Module):
    """Key layout of diffusers FeedForward: net.0.proj (GELU-tanh), net.2 (output Linear)."""

    def __init__(self, dim, dropout=0.0, activation_fn="gelu-approximate", final_dropout=False, inner_dim=None, bias=True):
        super().__init__()
        if activation_fn != "gelu-approximate":
            raise NotImplementedError("only activation_fn='gelu-approximate' (CogVideoX) is implemented")
        inner_dim = 4 * dim if inner_dim is None else inner_dim
        self.net = nn.ModuleList([_GELUProj(dim, inner_dim, bias), nn.Dropout(dropout), nn.Linear(inner_dim, dim, bias=bias)])
        if final_dropout:
            self.net.append(nn.Dropout(dropout))


# =================================================================================================== attention
class CogVideoXAttnProcessor2_0:
    """Plain processor (T2To stage / use_vip False).  Same call signature as attention_processor.py:1895-1902."""

    def __call__(self, attn: "Attention", hidden_states, encoder_hidden_states, attention_mask=None, image_rotary_emb=None):
        if attention_mask is not None:
            raise NotImplementedError("attention_mask is not used on the CogVideoX path")
        return _run_attention(attn, None, hidden_states, encoder_hidden_states, image_rotary_emb, None, None)


class VideoIPAdapterCogVideoXAttnProcessor2_0(nn.Module):
    """func_type "1" video-IP-adapter processor.  Owns vip_to_{q,k,v} / vip_norm_{q,k}; `scale` is set by the pipeline on
    modules whose class NAME equals this one (pipeline_cogvideox_mp_fifo.py:981-983)."""

    def __init__(self, heads, cross_attention_dim=None, dim_head=None, eps=1e-6, scale=1.0, qk_norm=None, bias=False,
                 num_tokens=None, **kwargs):
        super().__init__()
        hidden_size = heads * dim_head
        self.cross_attention_dim, self.scale, self.num_tokens = cross_attention_dim, scale, num_tokens
        self.vip_to_q = nn.Linear(cross_attention_dim, hidden_size, bias=bias)
        self.vip_to_k = nn.Linear(cross_attention_dim, hidden_size, bias=bias)
        self.vip_to_v = nn.Linear(cross_attention_dim, hidden_size, bias=bias)
        if qk_norm == "layer_norm":
            self.vip_norm_q = nn.LayerNorm(dim_head, eps=eps)
            self.vip_norm_k = nn.LayerNorm(dim_head, eps=eps)
        else:
            self.vip_norm_q = self.vip_norm_k = None

    def __call__(self, attn: "Attention", hidden_states, encoder_hidden_states, attention_mask=None, image_rotary_emb=None,
                 vip_image_rotary_emb=None, vip_condition_rotary_emb=None):
        if attention_mask is not None:
            raise NotImplementedError("attention_mask is not used on the CogVideoX path")
        return _run_attention(attn, self, hidden_states, encoder_hidden_states, image_rotary_emb, vip_image_rotary_emb,
                              vip_condition_rotary_emb)

    forward = __call__


class Attention(nn.Module):
    def __init__(self, query_dim, dim_head=64, heads=8, qk_norm=None, eps=1e-5, bias=False, out_bias=True, processor=None,
                 **_unused):
        super().__init__()
        self.inner_dim, self.heads, self.is_cross_attention = dim_head * heads, heads, False
        self.to_q = nn.Linear(query_dim, self.inner_dim, bias=bias)
        self.to_k = nn.Linear(query_dim, self.inner_dim, bias=bias)
        self.to_v = nn.Linear(query_dim, self.inner_dim, bias=bias)
        if qk_norm == "layer_norm":
            self.norm_q = nn.LayerNorm(dim_head, eps=eps)
            self.norm_k = nn.LayerNorm(dim_head, eps=eps)
        else:
            self.norm_q = self.norm_k = None
        self.to_out = nn.ModuleList([nn.Linear(self.inner_dim, query_dim, bias=out_bias), nn.Dropout(0.0)])
        self.set_processor(processor if processor is not None else CogVideoXAttnProcessor2_0())

    def set_processor(self, processor) -> None:
        # attention_processor.py:423-441: a module processor replaces a module processor in _modules
        if hasattr(self, "processor") and isinstance(self.processor, nn.Module) and not isinstance(processor, nn.Module):
            self._modules.pop("processor")
        self.processor = processor

    def get_processor(self):
        return self.processor

    def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None, **cross_attention_kwargs):
        # attention_processor.py:484-493: kwargs are filtered by the processor's __call__ parameter names
        names = set(inspect.signature(self.processor.__call__).parameters.keys())
        kw = {k: v for k, v in cross_attention_kwargs.items() if k in names}
        return self.processor(self, hidden_states, encoder_hidden_states=encoder_hidden_states,
                              attention_mask=attention_mask, **kw)


class _PackCache:
    """Concatenated projection weights, rebuilt when any source parameter is re-assigned or modified in place."""

    def __init__(self):
        self.key, self.w, self.b = None, None, None

    def get(self, linears):
        key = tuple((l.weight.data_ptr(), l.weight._version, None if l.bias is None else l.bias._version) for l in linears)
        if key != self.key:
            self.w = torch.cat([l.weight.detach() for l in linears], dim=0).contiguous()
            self.b = None if linears[0].bias is None else torch.cat([l.bias.detach() for l in linears]).contiguous()
            self.key = key
        return self.w, self.b


def _qkv_projs(attn, proc, outs, rows_tv, rope, img_rope, cond_rope):
    """tg_qkv_proj descriptors in the weight order [to_q,to_k,to_v,(vip_to_q,vip_to_k,vip_to_v)]."""
    projs = []

    def mk(out, out_rows, norm, video_rope, vip_rope):
        p = E.QkvProj()
        p.out, p.out_rows = out.data_ptr(), out_rows
        if norm is not None:
            p.ln_w, p.ln_b = norm.weight.data_ptr(), norm.bias.data_ptr()
        if video_rope is not None:
            p.cos_video, p.sin_video = video_rope[0].data_ptr(), video_rope[1].data_ptr()
        if vip_rope is not None:
            p.cos_vip, p.sin_vip = vip_rope[0].data_ptr(), vip_rope[1].data_ptr()
        return p

    projs.append(mk(outs[0], rows_tv, attn.norm_q, rope, None))
    projs.append(mk(outs[1], rows_tv, attn.norm_k, rope, None))
    projs.append(mk(outs[2], rows_tv, None, None, None))
    if proc is not None:
        rows_all = outs[3].shape[2]
        projs.append(mk(outs[3], rows_all, proc.vip_norm_q, img_rope, cond_rope))
        projs.append(mk(outs[4], rows_all, proc.vip_norm_k, img_rope, cond_rope))
        projs.append(mk(outs[5], rows_all, None, None, None))
    return projs


def _attention_core(attn, proc, y, B, rowmap, rope, img_rope, cond_rope, bufs):
    """y: normalised stream [B*rows, d] ordered [text|video|vip].  Fills bufs.A [B, rows, d] with the merged-head
    attention outputs (before to_out).  Kernel sequence K3 -> K4 -> K5 -> K6."""
    H = attn.heads
    n_tv = rowmap.n_text + rowmap.n_video
    linears = [attn.to_q, attn.to_k, attn.to_v] + ([proc.vip_to_q, proc.vip_to_k, proc.vip_to_v] if proc is not None else [])
    cache = attn.__dict__.setdefault("_tg_pack", _PackCache())
    w, b = cache.get(linears)
    eps = attn.norm_q.eps if attn.norm_q is not None else 1e-6
    projs = _qkv_projs(attn, proc, bufs.qkv, n_tv, rope, img_rope, cond_rope)
    if getattr(bufs, "sp", None) is not None:
        return _attention_core_sp(attn, proc, y, w, b, eps, projs, B, rowmap, bufs)
    E.qkv_rope_gemm(y, w, b, B, H, rowmap, projs, eps)
    qb, kb, vb = bufs.qkv[:3]
    side = None
    if proc is not None and _SIDE_STREAM:
        # K6 (480 vip queries x 18 256 keys: 192 CTAs, 1.3 waves on 148 SMs) is independent of K4/K5 — it writes other rows
        # of A — so it runs on a side stream and its CTAs fill the SMs the self-attention's last wave leaves idle.
        qv, kv, vv = bufs.qkv[3:]
        main = torch.cuda.current_stream()
        side = bufs.side_stream
        side.wait_stream(main)
        with torch.cuda.stream(side):
            E.attn_fwd(qv, kv, vv, bufs.A, q_row0=n_tv, q_rows=rowmap.n_vip, out_row0=n_tv)
    scales = None
    if proc is not None:
        scale = proc.scale
        scales = [float(s) for s in (scale if isinstance(scale, (list, tuple)) else [scale])]
        if len(scales) != B:
            scales = [scales[0]] * B  # attention_processor.py:2130-2131
    if scales is not None and len(set(scales)) == 1 and _FUSE_PAIR:
        qv, kv, vv = bufs.qkv[3:]   # K4 + K5 in one launch
        E.attn_fwd_pair(qb, kb, vb, n_tv, n_tv, qv, kv, vv, n_tv, rowmap.n_vip, bufs.A, _bf16_scalar(scales[0]))
    else:
        E.attn_fwd(qb, kb, vb, bufs.A, out_row0=0)
    if proc is not None:
        qv, kv, vv = bufs.qkv[3:]
        if len(set(scales)) == 1 and _FUSE_PAIR:
            pass
        elif len(set(scales)) == 1:
            E.attn_fwd(qv, kv, vv, bufs.A, q_row0=0, q_rows=n_tv, kv_row0=n_tv, kv_rows=rowmap.n_vip, out_row0=0,
                       accumulate=True, out_scale=_bf16_scalar(scales[0]))
        else:
            for bi, s in enumerate(scales):
                E.attn_fwd(qv[bi:bi + 1], kv[bi:bi + 1], vv[bi:bi + 1], bufs.A[bi:bi + 1], q_row0=0, q_rows=n_tv,
                           kv_row0=n_tv, kv_rows=rowmap.n_vip, out_row0=0, accumulate=True, out_scale=_bf16_scalar(s))
        if side is None:
            E.attn_fwd(qv, kv, vv, bufs.A, q_row0=n_tv, q_rows=rowmap.n_vip, out_row0=n_tv)
        else:
            torch.cuda.current_stream().wait_stream(side)


def _attention_core_sp(attn, proc, y, w, b, eps, projs, B, rowmap, bufs):
    """Sequence-parallel form of the kernel sequence K3 -> K4 -> K5 -> K6 (seqpar.py): y holds this rank's rows; the Q/K/V
    GEMM scatters heads to their owner ranks, the attention kernels (this rank's heads, all rows) scatter output rows to
    theirs.  Two cross-rank stream barriers order producers before consumers; the second also keeps the next layer's Q/K/V
    stores behind every rank's attention reads."""
    sp = bufs.sp
    n_tv = rowmap.n_text + rowmap.n_video
    E.qkv_rope_gemm(y, w, b, B, attn.heads, rowmap, projs, eps, scatter=bufs.qkv_scatter)
    sp.barrier()
    qb, kb, vb = bufs.qkv[:3]
    out = bufs.attn_scatter
    if proc is None:
        E.attn_fwd(qb, kb, vb, out, out_row0=0)
    else:
        qv, kv, vv = bufs.qkv[3:]
        # K6 (480 vip queries x all keys) has only 2 * B * H/world CTAs, each walking every key block: on a side stream its
        # CTAs share the SMs with K4 + K5 instead of holding a few of them alone (it writes other rows of the output)
        main, side = torch.cuda.current_stream(), bufs.side_stream
        side.wait_stream(main)
        with torch.cuda.stream(side):
            E.attn_fwd(qv, kv, vv, out, q_row0=n_tv, q_rows=rowmap.n_vip, out_row0=n_tv)
        scale = proc.scale
        scales = [float(s_) for s_ in (scale if isinstance(scale, (list, tuple)) else [scale])]
        if len(scales) != B:
            scales = [scales[0]] * B  # attention_processor.py:2130-2131
        if len(set(scales)) == 1:
            # K4 + K5 in one launch: the vip cross-attention is added before the row leaves the SM, so the peer buffer is
            # written once instead of read-modified-written over NVLink
            E.attn_fwd_pair(qb, kb, vb, n_tv, n_tv, qv, kv, vv, n_tv, rowmap.n_vip, out, _bf16_scalar(scales[0]))
        else:
            for bi, s_ in enumerate(scales):
                E.attn_fwd_pair(qb[bi:bi + 1], kb[bi:bi + 1], vb[bi:bi + 1], n_tv, n_tv, qv[bi:bi + 1], kv[bi:bi + 1],
                                vv[bi:bi + 1], n_tv, rowmap.n_vip, bufs.attn_scatter_for_batch(bi), _bf16_scalar(s_))
        main.wait_stream(side)
    sp.barrier()


# Round 1 measured both neutral on the power-capped step (839.8 ms either way, profiles/r01_energy.md).  With the round-2
# attention kernel (speculative reference, 1/8 of the exponentials on the FMA pipe) they pay: 829.5 ms with three launches,
# 813.6 with K4 + K5 in one launch, 806.0 with K6 on the side stream as well (profiles/r02_ab_bench_2.jsonl) — both on.
# The fused pair is also what the sequence-parallel forward runs, so sharded and unsharded forwards stay bit-identical.
_SIDE_STREAM = os.environ.get("TG_SIDE_STREAM", "1") != "0"
_FUSE_PAIR = os.environ.get("TG_FUSE_PAIR", "1") != "0"


class _Buffers:
    """Workspaces of one (B, rows, d, H, vip) geometry; allocated once and reused by every layer and every step."""

    def __init__(self, B, rowmap, d, H, ff_dim, use_vip, device):
        rows, n_tv = rowmap.rows_per_batch, rowmap.n_text + rowmap.n_video
        bf = dict(device=device, dtype=torch.bfloat16)
        self.X = torch.empty(B, rows, d, **bf)
        self.Y = torch.empty(B, rows, d, **bf)
        self.A = torch.empty(B, rows, d, **bf)
        self.Hff = torch.empty(B * rows, ff_dim, **bf) if ff_dim else None
        self.qkv = [torch.empty(B, H, n_tv, 64, **bf) for _ in range(3)]
        if use_vip:
            self.qkv += [torch.empty(B, H, rows, 64, **bf) for _ in range(3)]
        self.side_stream = torch.cuda.Stream(device=device) if torch.device(device).type == "cuda" else None


def _run_attention(attn, proc, hidden_states, encoder_hidden_states, rope, img_rope, cond_rope):
    """Module-level entry (the diffusers processor contract): inputs are the already-normalised streams."""
    if not hidden_states.is_cuda:
        raise E.TokensGenError("tokensgen_b200 attention processors run on CUDA only (no CPU fallback)")
    B, n_video, d = hidden_states.shape
    n_vip = proc.num_tokens if proc is not None else 0
    n_text = encoder_hidden_states.shape[1] - n_vip
    dev = hidden_states.device
    rowmap = E.make_rowmap(n_text, n_video, n_vip, n_video, 1)  # RoPE/LN here do not need the frame split
    key = (B, n_text, n_video, n_vip, d, str(dev))
    cache = attn.__dict__.setdefault("_tg_bufs", {})
    if key not in cache:
        cache.clear()
        cache[key] = _Buffers(B, rowmap, d, attn.heads, 0, proc is not None, dev)
    bufs = cache[key]
    enc = encoder_hidden_states.to(torch.bfloat16)
    parts = [enc[:, :n_text], hidden_states.to(torch.bfloat16)] + ([enc[:, n_text:]] if n_vip else [])
    torch.cat(parts, dim=1, out=bufs.Y)
    _attention_core(attn, proc, bufs.Y.view(B * rowmap.rows_per_batch, d), B, rowmap, _rope_pair(rope, dev),
                    _rope_pair(img_rope, dev), _rope_pair(cond_rope, dev), bufs)
    lin = attn.to_out[0]
    out = E.gemm_bias_act(bufs.A.view(-1, d), lin.weight, lin.bias).view(B, rowmap.rows_per_batch, -1)
    text_o, hid_o, vip_o = out.split([n_text, n_video, n_vip], dim=1)
    return hid_o, (torch.cat([text_o, vip_o], dim=1) if n_vip else text_o)


# =================================================================================================== block
class CogVideoXBlock(nn.Module):
    def __init__(self, dim, num_attention_heads, attention_head_dim, time_embed_dim, dropout=0.0,
                 activation_fn="gelu-approximate", attention_bias=False, qk_norm=True, norm_elementwise_affine=True,
                 norm_eps=1e-5, final_dropout=True, ff_inner_dim=None, ff_bias=True, attention_out_bias=True):
        super().__init__()
        self.use_vip = False
        self.time_embed_dim, self.dim, self.norm_elementwise_affine, self.norm_eps = \
            time_embed_dim, dim, norm_elementwise_affine, norm_eps
        self.attention_head_dim, self.num_attention_heads = attention_head_dim, num_attention_heads
        self.qk_norm, self.attention_bias, self.attention_out_bias = qk_norm, attention_bias, attention_out_bias
        self.norm1 = CogVideoXLayerNormZero(time_embed_dim, dim, norm_elementwise_affine, norm_eps, bias=True)
        self.attn1 = Attention(query_dim=dim, dim_head=attention_head_dim, heads=num_attention_heads,
                               qk_norm="layer_norm" if qk_norm else None, eps=1e-6, bias=attention_bias,
                               out_bias=attention_out_bias, processor=CogVideoXAttnProcessor2_0())
        self.norm2 = CogVideoXLayerNormZero(time_embed_dim, dim, norm_elementwise_affine, norm_eps, bias=True)
        self.ff = FeedForward(dim, dropout=dropout, activation_fn=activation_fn, final_dropout=final_dropout,
                              inner_dim=ff_inner_dim, bias=ff_bias)

    def set_vip_layers(self, **kwargs):
        """cogvideox_transformer_3d.py:145-218 for func_type "1" (the shipped one)."""
        self.use_vip = True
        self.vip_length = kwargs["length"]
        self.vip_func_type = kwargs["func_type"]
        if self.vip_func_type != "1":
            raise NotImplementedError("only video-IP-adapter func_type '1' is on the reproduced path (config/infer/*.yaml)")
        self.vip_norm1 = CogVideoXVIPLayerNormZero(self.time_embed_dim, self.dim, self.norm_elementwise_affine, self.norm_eps, bias=True)
        self.vip_norm2 = CogVideoXVIPLayerNormZero(self.time_embed_dim, self.dim, self.norm_elementwise_affine, self.norm_eps, bias=True)
        proc = VideoIPAdapterCogVideoXAttnProcessor2_0(
            dim_head=self.attention_head_dim, heads=self.num_attention_heads, cross_attention_dim=self.dim, eps=1e-6,
            scale=kwargs["scale"], qk_norm="layer_norm" if self.qk_norm else None, bias=self.attention_bias,
            num_tokens=kwargs["length"])
        proc.to(device=self.attn1.to_q.weight.device, dtype=self.attn1.to_q.weight.dtype)
        self.attn1.set_processor(proc)
        with torch.no_grad():  # :207-218 — vip projections start as copies of the base ones
            for a, b in (("vip_to_q", "to_q"), ("vip_to_k", "to_k"), ("vip_to_v", "to_v")):
                getattr(proc, a).weight.copy_(getattr(self.attn1, b).weight)
                if self.attention_bias:
                    getattr(proc, a).bias.copy_(getattr(self.attn1, b).bias)
            if self.qk_norm:
                for a, b in (("vip_norm_q", "norm_q"), ("vip_norm_k", "norm_k")):
                    getattr(proc, a).weight.copy_(getattr(self.attn1, b).weight)
                    getattr(proc, a).bias.copy_(getattr(self.attn1, b).bias)

    # ---- engine pieces shared with the fused model forward
    def _ada_linears(self):
        mods = [self.norm1, self.norm2] + ([self.vip_norm1, self.vip_norm2] if self.use_vip else [])
        return [m.linear for m in mods]

    def _run(self, bufs: _Buffers, B: int, rowmap, ada: torch.Tensor, col0: int, rope, img_rope, cond_rope):
        """One block on the resident stream bufs.X.  `ada` [B*frames, *]: this block's AdaLN vectors start at column col0
        in the order norm1(6d) | norm2(6d) | vip_norm1(3d) | vip_norm2(3d)."""
        d = self.dim
        rows = E.rows_local(rowmap)  # all rows, or this rank's shard (sequence parallel)
        installed = self.attn1.processor
        if type(installed) not in (CogVideoXAttnProcessor2_0, VideoIPAdapterCogVideoXAttnProcessor2_0) or \
                isinstance(installed, VideoIPAdapterCogVideoXAttnProcessor2_0) != self.use_vip:
            # the fused engine runs the two built-in processors' kernel sequence on the resident stream; a processor installed
            # with Attention.set_processor is honoured by Attention.forward (the plugin API), never silently replaced here
            raise E.TokensGenError(
                f"attn1.processor is {type(installed).__name__}: the fused model forward only executes the built-in "
                "CogVideoXAttnProcessor2_0 / VideoIPAdapterCogVideoXAttnProcessor2_0; call the block's modules through "
                "Attention.forward to run a custom processor")
        proc = installed if self.use_vip else None
        X2, Y2 = bufs.X.view(B * rows, d), bufs.Y.view(B * rows, d)
        c = lambda i: ada[:, col0 + i * d: col0 + (i + 1) * d]
        vip = self.use_vip and rowmap.n_vip > 0
        # norm1 chunks: shift0 scale1 gate2 enc_shift3 enc_scale4 enc_gate5; vip_norm1 at 12..14: shift scale gate
        shift = E.make_modvec(c(3), c(0), c(12) if vip else None)
        scale = E.make_modvec(c(4), c(1), c(13) if vip else None)
        gate = E.make_modvec(c(5), c(2), c(14) if vip else None)
        vn1 = self.vip_norm1.norm if vip else None
        E.ln_modulate(X2, Y2, B, rowmap, self.norm1.norm.weight, self.norm1.norm.bias,
                      vn1.weight if vip else None, vn1.bias if vip else None, self.norm_eps, shift, scale)
        _attention_core(self.attn1, proc, Y2, B, rowmap, rope, img_rope, cond_rope, bufs)
        lin = self.attn1.to_out[0]
        E.gemm_gate_residual(bufs.A.view(B * rows, d), lin.weight, lin.bias, X2, B, rowmap, gate)
        # norm2 chunks at 6..11; vip_norm2 at 15..17
        shift = E.make_modvec(c(9), c(6), c(15) if vip else None)
        scale = E.make_modvec(c(10), c(7), c(16) if vip else None)
        gate = E.make_modvec(c(11), c(8), c(17) if vip else None)
        vn2 = self.vip_norm2.norm if vip else None
        E.ln_modulate(X2, Y2, B, rowmap, self.norm2.norm.weight, self.norm2.norm.bias,
                      vn2.weight if vip else None, vn2.bias if vip else None, self.norm_eps, shift, scale)
        ff1, ff2 = self.ff.net[0].proj, self.ff.net[2]
        E.gemm_bias_act(Y2, ff1.weight, ff1.bias, bufs.Hff, act=E.ACT_GELU_TANH)
        E.gemm_gate_residual(bufs.Hff, ff2.weight, ff2.bias, X2, B, rowmap, gate)

    def forward(self, hidden_states, encoder_hidden_states, temb, image_rotary_emb=None, vip_image_rotary_emb=None,
                vip_condition_rotary_emb=None, attention_mask=None):
        """Standalone block call with the reference's signature (cogvideox_transformer_3d.py:221-230)."""
        if not hidden_states.is_cuda:
            raise E.TokensGenError("tokensgen_b200 blocks run on CUDA only (no CPU fallback)")
        B, n_video, d = hidden_states.shape
        frames = temb.shape[1]
        n_vip = self.vip_length if self.use_vip else 0
        n_text = encoder_hidden_states.shape[1] - n_vip
        dev = hidden_states.device
        rowmap = E.make_rowmap(n_text, n_video, n_vip, n_video // frames, frames)
        key = (B, n_text, n_video, n_vip, frames, str(dev))
        cache = self.__dict__.setdefault("_tg_bufs", {})
        if key not in cache:
            cache.clear()
            cache[key] = _Buffers(B, rowmap, d, self.num_attention_heads, self.ff.net[0].proj.out_features, self.use_vip, dev)
        bufs = cache[key]
        enc = encoder_hidden_states.to(torch.bfloat16)
        parts = [enc[:, :n_text], hidden_states.to(torch.bfloat16)] + ([enc[:, n_text:]] if n_vip else [])
        torch.cat(parts, dim=1, out=bufs.X)
        packc = self.__dict__.setdefault("_tg_ada_pack", _PackCache())
        w, b = packc.get(self._ada_linears())
        silu_t = torch.nn.functional.silu(temb.to(torch.bfloat16)).reshape(B * frames, -1).contiguous()
        ada = E.gemm_bias_act(silu_t, w, b)
        self._run(bufs, B, rowmap, ada, 0, _rope_pair(image_rotary_emb, dev), _rope_pair(vip_image_rotary_emb, dev),
                  _rope_pair(vip_condition_rotary_emb, dev))
        text_o, hid_o, vip_o = bufs.X.split([n_text, n_video, n_vip], dim=1)
        return hid_o.clone(), (torch.cat([text_o, vip_o], dim=1) if n_vip else text_o.clone())


# =================================================================================================== model
class CogVideoXTransformer3DModel(nn.Module):
    """Same constructor / forward signature / state-dict layout as cogvideox_transformer_3d.py:335-770 (RoPE models)."""

    def __init__(self, num_attention_heads=30, attention_head_dim=64, in_channels=16, out_channels=16, flip_sin_to_cos=True,
                 freq_shift=0, time_embed_dim=512, text_embed_dim=4096, num_layers=30, dropout=0.0, attention_bias=True,
                 sample_width=90, sample_height=60, sample_frames=49, patch_size=2, temporal_compression_ratio=4,
                 max_text_seq_length=226, activation_fn="gelu-approximate", timestep_activation_fn="silu",
                 norm_elementwise_affine=True, norm_eps=1e-5, spatial_interpolation_scale=1.875,
                 temporal_interpolation_scale=1.0, use_rotary_positional_embeddings=False,
                 use_learned_positional_embeddings=False, use_output_projection=True):
        super().__init__()
        cfg = {k: v for k, v in locals().items() if k not in ("self", "__class__")}
        self.config = SimpleNamespace(**cfg)
        if not use_rotary_positional_embeddings:
            raise NotImplementedError("only the CogVideoX-5b layout (use_rotary_positional_embeddings=True) is implemented")
        if not use_output_projection:
            raise NotImplementedError("use_output_projection=False is not on the reproduced path")
        if attention_head_dim != 64:
            raise NotImplementedError("the attention kernels are specialised for attention_head_dim=64")
        inner_dim = num_attention_heads * attention_head_dim
        self.use_vip = False
        self.patch_embed = CogVideoXPatchEmbed(patch_size=patch_size, in_channels=in_channels, embed_dim=inner_dim,
                                               text_embed_dim=text_embed_dim, bias=True)
        self.time_proj = Timesteps(inner_dim, flip_sin_to_cos, freq_shift)
        self.time_embedding = TimestepEmbedding(inner_dim, time_embed_dim, timestep_activation_fn)
        self.transformer_blocks = nn.ModuleList([
            CogVideoXBlock(dim=inner_dim, num_attention_heads=num_attention_heads, attention_head_dim=attention_head_dim,
                           time_embed_dim=time_embed_dim, dropout=dropout, activation_fn=activation_fn,
                           attention_bias=attention_bias, norm_elementwise_affine=norm_elementwise_affine, norm_eps=norm_eps)
            for _ in range(num_layers)])
        self.norm_final = nn.LayerNorm(inner_dim, norm_eps, norm_elementwise_affine)
        self.norm_out = AdaLayerNorm(embedding_dim=time_embed_dim, output_dim=2 * inner_dim,
                                     norm_elementwise_affine=norm_elementwise_affine, norm_eps=norm_eps, chunk_dim=1)
        self.proj_out = nn.Linear(inner_dim, patch_size * patch_size * out_channels)
        self._tg_bufs: Dict[Any, _Buffers] = {}
        self._tg_ada = _PackCache()

    # ---- reference API
    @property
    def dtype(self):
        return self.proj_out.weight.dtype

    @property
    def device(self):
        return self.proj_out.weight.device

    @property
    def attn_processors(self):
        return {f"transformer_blocks.{i}.attn1.processor": b.attn1.get_processor()
                for i, b in enumerate(self.transformer_blocks)}

    def set_vip_layers(self, vip_ckpt_dir=None, **kwargs):
        """cogvideox_transformer_3d.py:591-622."""
        self.use_vip = True
        self.vip_length = kwargs["length"]
        self.vip_func_type = kwargs["func_type"]
        self.patch_embed.set_vip_layers(**kwargs)
        self.patch_embed.vip_proj.to(device=self.device, dtype=self.dtype)
        for block in self.transformer_blocks:
            block.set_vip_layers(**kwargs)
            block.vip_norm1.to(device=self.device, dtype=self.dtype)
            block.vip_norm2.to(device=self.device, dtype=self.dtype)
        if vip_ckpt_dir is not None:
            path = os.path.join(vip_ckpt_dir, "vip.pt")
            if not os.path.exists(path):
                raise IOError(f"no vip weights found in {vip_ckpt_dir}")
            sd = torch.load(path, weights_only=True)
            own = self.state_dict().keys()
            for key in sd.keys():
                assert key in own, key
            self.load_state_dict(sd, strict=False)

    def enable_sequence_parallel(self, group=None) -> None:
        """Runs every later forward sequence-parallel (Ulysses) over `group` (default: all ranks): all ranks must call
        forward with identical inputs and all get the full output.  See tokensgen_b200/seqpar.py."""
        import torch.distributed as dist
        from .seqpar import SeqParallel
        # one SeqParallel (and, below, one set of peer-mapped workspaces per geometry) per group for the life of the model:
        # the pipelines switch this on and off around every base stage, and a symmetric-memory segment + rendezvous per
        # video would never be freed
        cache = self.__dict__.setdefault("_tg_sp_cache", {})
        g = group if group is not None else dist.group.WORLD
        if g not in cache:
            cache[g] = SeqParallel(group)
        self.__dict__["_tg_sp"] = cache[g]

    def disable_sequence_parallel(self) -> None:
        self.__dict__.pop("_tg_sp", None)

    def save_vip_layers(self, vip_ckpt_dir=None):
        assert self.use_vip
        out = {n: p.to("cpu").to(torch.float32) for n, p in self.named_parameters() if "vip_" in n}
        os.makedirs(vip_ckpt_dir, exist_ok=True)
        torch.save(out, os.path.join(vip_ckpt_dir, "vip.pt"))

    # ---- forward
    def _padded(self, name: str, param: torch.Tensor, fn):
        """Zero-padded copy of a narrow weight, rebuilt when the parameter is re-assigned or modified in place."""
        cache = self.__dict__.setdefault("_tg_padded", {})
        key = (param.data_ptr(), param._version)
        if name not in cache or cache[name][0] != key:
            cache[name] = (key, fn(param.detach()).contiguous())
        return cache[name][1]

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, subfolder=None, torch_dtype=None, **kwargs):
        """diffusers-style loader (infer_cogvideo_mp_fifo.py:150-156): <path>/<subfolder>/config.json + weights."""
        from .loading import build_from_pretrained
        kwargs.pop("revision", None), kwargs.pop("variant", None)
        return build_from_pretrained(cls, pretrained_model_name_or_path, subfolder, torch_dtype, **kwargs)

    def _ada_table(self, silu_emb: torch.Tensor) -> torch.Tensor:
        """K1: every AdaLN linear of every block (and norm_out) in ONE GEMM — temb does not depend on the layer."""
        linears = [l for blk in self.transformer_blocks for l in blk._ada_linears()] + [self.norm_out.linear]
        w, b = self._tg_ada.get(linears)
        return E.gemm_bias_act(silu_emb, w, b)

    def forward(self, hidden_states, encoder_hidden_states, timestep, vip_encoder_hidden_states=None, timestep_cond=None,
                image_rotary_emb=None, vip_image_rotary_emb=None, vip_condition_rotary_emb=None, vip_grid_t=None,
                attention_kwargs=None, return_dict: bool = True):
        if not hidden_states.is_cuda:
            raise E.TokensGenError("CogVideoXTransformer3DModel (tokensgen_b200) runs on CUDA only (no CPU fallback)")
        if self.dtype != torch.bfloat16:
            raise E.TokensGenError("tokensgen_b200 computes in bf16: call model.to(torch.bfloat16)")
        if attention_kwargs is not None and attention_kwargs.get("attention_masks") is not None:
            raise NotImplementedError("attention masks are not used on the CogVideoX path")
        cfg = self.config
        dev = hidden_states.device
        B, F, C, Hh, Ww = hidden_states.shape
        p = cfg.patch_size
        d = cfg.num_attention_heads * cfg.attention_head_dim
        hw = (Hh // p) * (Ww // p)
        n_text = encoder_hidden_states.shape[1]
        n_video = F * hw
        use_vip = self.use_vip and vip_encoder_hidden_states is not None
        if self.use_vip and not use_vip:
            raise E.TokensGenError("set_vip_layers() was called: vip_encoder_hidden_states is required")
        n_vip = self.vip_length if use_vip else 0

        # 1. time embedding (K10) + all AdaLN vectors (K1)
        ts = timestep.reshape(-1) if torch.is_tensor(timestep) else torch.tensor([timestep] * B)
        frames = ts.numel() // B  # F when per-frame timesteps are given, else 1
        te = self.time_embedding
        _, silu_emb = E.time_embedding(ts, te.linear_1.weight, te.linear_1.bias, te.linear_2.weight, te.linear_2.bias, d,
                                       cfg.flip_sin_to_cos, float(cfg.freq_shift))
        ada = self._ada_table(silu_emb)

        rowmap = E.make_rowmap(n_text, n_video, n_vip, n_video // frames, frames)
        sp = self.__dict__.get("_tg_sp")
        # `frames` is part of the key: the sharded workspaces freeze their rowmap (frames / hw decide which AdaLN row a video
        # row reads), so a per-frame-timestep forward must never reuse the buffers of a per-sample one
        key = (B, n_text, n_video, n_vip, frames, str(dev))
        ff_dim = self.transformer_blocks[0].ff.net[0].proj.out_features
        if sp is None:
            if key not in self._tg_bufs:
                self._tg_bufs.clear()
                self._tg_bufs[key] = _Buffers(B, rowmap, d, cfg.num_attention_heads, ff_dim, use_vip, dev)
            bufs = self._tg_bufs[key]
        else:
            from .seqpar import ShardedBuffers
            spc = self.__dict__.setdefault("_tg_bufs_sp", {})
            skey = key + (id(sp),)
            if skey not in spc:
                while len(spc) >= 6:          # base stage / T2To geometries and the FIFO ramp's 2-, 4-, 8-rank groups; older ones are dropped
                    spc.pop(next(iter(spc)))
                spc[skey] = ShardedBuffers(sp, B, rowmap, d, cfg.num_attention_heads, ff_dim, use_vip, dev)
            bufs = spc[skey]
        x_full = bufs.X if sp is None else bufs.Xfull  # sequence parallel: every rank embeds all rows, keeps its shard

        # 2. patch embedding (K9) straight into the resident stream [text | video | vip]
        pe = self.patch_embed
        patches = E.patchify(hidden_states.to(torch.bfloat16).contiguous(), p)
        text = encoder_hidden_states.to(torch.bfloat16)
        pw = pe.proj.weight.reshape(d, -1)
        if pw.shape[1] % 64:  # patch_size 1 (the T2To model, train_cogvideo_t2to.py:1277): K = 16 -> zero-pad to the GEMM's 64
            pw = self._padded("patch_w", pe.proj.weight, lambda w: torch.nn.functional.pad(w.reshape(d, -1), (0, -w[0].numel() % 64)))
            patches = torch.nn.functional.pad(patches, (0, pw.shape[1] - patches.shape[1]))
        if use_vip:
            vip_rows = vip_encoder_hidden_states.to(torch.bfloat16).permute(0, 1, 3, 4, 2).reshape(B, n_vip, -1).contiguous()
        for b in range(B):
            E.gemm_bias_act(text[b].contiguous(), pe.text_proj.weight, pe.text_proj.bias, x_full[b, :n_text])
            E.gemm_bias_act(patches[b * n_video:(b + 1) * n_video], pw, pe.proj.bias, x_full[b, n_text:n_text + n_video])
            if use_vip:
                E.gemm_bias_act(vip_rows[b], pe.vip_proj.weight, pe.vip_proj.bias, x_full[b, n_text + n_video:])
        if sp is not None:
            bufs.X.copy_(x_full[:, bufs.row0:bufs.row0 + bufs.rows_local])
            rowmap = bufs.rowmap  # same layout + this rank's (row0, rows_local)

        # 3. blocks
        rope = _rope_pair(image_rotary_emb, dev)
        img_rope = _rope_pair(vip_image_rotary_emb, dev) if use_vip else None
        cond_rope = _rope_pair(vip_condition_rotary_emb, dev) if use_vip else None
        per_block = (18 if self.use_vip else 12) * d
        for i, blk in enumerate(self.transformer_blocks):
            blk._run(bufs, B, rowmap, ada, i * per_block, rope, img_rope, cond_rope)

        # 4. norm_final -> norm_out (shift, scale order) -> proj_out -> unpatchify (K11)
        col = len(self.transformer_blocks) * per_block
        shift = E.make_modvec(None, ada[:, col:col + d], None)
        scale = E.make_modvec(None, ada[:, col + d:col + 2 * d], None)
        rows_l = E.rows_local(rowmap)
        X2, Y2 = bufs.X.view(B * rows_l, d), bufs.Y.view(B * rows_l, d)
        E.ln_modulate(X2, Y2, B, rowmap, self.norm_final.weight, self.norm_final.bias, None, None, cfg.norm_eps, shift, scale,
                      ln2_w=self.norm_out.norm.weight, ln2_b=self.norm_out.norm.bias, eps2=cfg.norm_eps)
        n_out = self.proj_out.out_features
        ow, ob = self.proj_out.weight, self.proj_out.bias
        if n_out % 64:  # patch_size 1: N = 16 -> zero rows up to 64, sliced away below
            ow = self._padded("out_w", self.proj_out.weight, lambda w: torch.nn.functional.pad(w, (0, 0, 0, -w.shape[0] % 64)))
            ob = self._padded("out_b", self.proj_out.bias, lambda b_: torch.nn.functional.pad(b_, (0, -b_.shape[0] % 64)))
        if sp is None:
            out_rows = torch.empty(B * n_video, ow.shape[0], device=dev, dtype=torch.bfloat16)
            for b in range(B):
                E.gemm_bias_act(bufs.Y[b, n_text:n_text + n_video], ow, ob, out_rows[b * n_video:(b + 1) * n_video])
        else:
            # proj_out over this rank's rows (text / vip rows of Y hold stale data: they are sliced away after the gather),
            # then ONE small all-gather of the [rows, p*p*C] projections
            loc = torch.zeros(B, bufs.chunk, ow.shape[0], device=dev, dtype=torch.bfloat16)
            for b in range(B):
                E.gemm_bias_act(bufs.Y[b], ow, ob, loc[b, :rows_l])
            out_rows = sp.gather_rows(loc, bufs.chunk, rowmap.rows_per_batch)[:, n_text:n_text + n_video]
            out_rows = out_rows.reshape(B * n_video, ow.shape[0]).contiguous()
        if ow.shape[0] != n_out:
            out_rows = out_rows[:, :n_out].contiguous()
        output = E.unpatchify(out_rows, B, F, self.proj_out.out_features // (p * p), Hh, Ww, p)
        if not return_dict:
            return (output,)
        return SimpleNamespace(sample=output)
