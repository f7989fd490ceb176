"""Host-side mirror of the reference's video-IP-adapter Resampler (longvgen/video_ipadapter/resampler.py:66-245), executing
on the C-ABI CUDA library.  Same class names, constructor arguments and state-dict keys (`latents`, `proj_in`, `proj_out`,
`norm_out`, `layers.N.0.{norm1,norm2,to_q,to_kv,to_out,norm_q,norm_k}`, `layers.N.1.net.{0.proj,2}`), same forward
signature `forward(x, image_rotary_emb, sampling_rotary_emb)`.

Kernel mapping (SURVEY K14 — no new kernel, the DiT ones are reused with a different geometry: 16 heads, 384 queries):
  norm1(x) / norm2(latents)        tg_ln_modulate with all-zero shift/scale tables, written straight into ONE [x; latents]
                                   buffer (the reference's torch.cat((x, latents)), resampler.py:101, never materialises)
  to_kv + norm_k + RoPE            tg_qkv_rope_gemm, 2 projections (k: LayerNorm(64) + image RoPE on the x rows and sampling
                                   RoPE on the latent rows, resampler.py:112-118; v: plain)
  to_q + norm_q + RoPE             tg_qkv_rope_gemm, 1 projection on the latent rows
  SDPA 384 x (n + 384)             tg_attn_fwd
  to_out + residual, FFN           tg_gemm_gate_residual with a unit gate / tg_gemm_bias_act (GELU-tanh)
The optional PCA truncation (resampler.py:230-237) is two plain fp32 matmuls in the PCA object's own dtype, as in the reference.
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Optional

import torch
import torch.nn as nn

from . import _ext as E
from .transformer import FeedForward, _PackCache, _rope_pair


class PerceiverAttention(nn.Module):
    def __init__(self, *, dim, dim_head=64, heads=8, qk_norm=True):
        super().__init__()
        if dim_head != 64:
            raise NotImplementedError("the attention kernels are specialised for dim_head=64")
        self.scale, self.dim_head, self.heads, self.qk_norm = dim_head ** -0.5, dim_head, heads, qk_norm
        inner = dim_head * heads
        self.norm1 = nn.LayerNorm(dim)
        self.norm2 = nn.LayerNorm(dim)
        self.to_q = nn.Linear(dim, inner, bias=False)
        self.to_kv = nn.Linear(dim, inner * 2, bias=False)
        self.to_out = nn.Linear(inner, dim, bias=False)
        if qk_norm:
            self.norm_q = nn.LayerNorm(dim_head, eps=1e-6)
            self.norm_k = nn.LayerNorm(dim_head, eps=1e-6)


class _Workspace:
    def __init__(self, B, n_x, l, dim, heads, ff_dim, device):
        bf = dict(device=device, dtype=torch.bfloat16)
        self.xp = torch.empty(B, n_x, dim, **bf)          # proj_in(x), constant across layers
        self.lat = torch.empty(B, l, dim, **bf)           # the latent stream, updated in place
        self.cat = torch.empty(B, n_x + l, dim, **bf)     # [norm1(x); norm2(latents)]
        self.q = torch.empty(B, heads, l, 64, **bf)
        self.k = torch.empty(B, heads, n_x + l, 64, **bf)
        self.v = torch.empty(B, heads, n_x + l, 64, **bf)
        self.att = torch.empty(B, l, heads * 64, **bf)
        self.ff = torch.empty(B * l, ff_dim, **bf)
        self.zeros = torch.zeros(1, dim, **bf)
        self.ones = torch.ones(1, dim, **bf)


class Resampler(nn.Module):
    def __init__(self, dim=1024, depth=8, dim_head=64, heads=16, num_height_queries=6, num_width_queries=6,
                 num_temporal_queries=13, embedding_dim=1280, output_dim=1024, ff_mult=4, max_height_seq_len: int = 16,
                 max_width_seq_len: int = 16, max_temporal_seq_len: int = 49, dropout: float = 0.0,
                 activation_fn: str = "gelu-approximate", ff_inner_dim: Optional[int] = None, final_dropout: bool = True,
                 ff_bias: bool = True, **kwargs):
        super().__init__()
        cfg = {k: v for k, v in locals().items() if k not in ("self", "__class__", "kwargs")}
        self.config = SimpleNamespace(**cfg)
        self.num_height_queries, self.num_width_queries, self.num_temporal_queries = \
            num_height_queries, num_width_queries, num_temporal_queries
        n_lat = num_height_queries * num_width_queries * num_temporal_queries
        self.latents = nn.Parameter(torch.randn(1, n_lat, dim) / dim ** 0.5)
        self.proj_in = nn.Linear(embedding_dim, dim)
        self.proj_out = nn.Linear(dim, output_dim)
        self.norm_out = nn.LayerNorm(output_dim)
        self.pca = None
        self.layers = nn.ModuleList([
            nn.ModuleList([PerceiverAttention(dim=dim, dim_head=dim_head, heads=heads, qk_norm=True),
                           FeedForward(dim=dim, dropout=dropout, activation_fn=activation_fn, final_dropout=final_dropout,
                                       inner_dim=ff_inner_dim, bias=ff_bias)])
            for _ in range(depth)])
        self._tg_ws = {}

    @property
    def dtype(self):
        return self.proj_in.weight.dtype

    @property
    def device(self):
        return self.proj_in.weight.device

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, subfolder=None, torch_dtype=None, **kwargs):
        """diffusers-style loader (infer_cogvideo_mp_fifo.py:162-166): <path>/<subfolder>/config.json + weights."""
        from .loading import build_from_pretrained
        return build_from_pretrained(cls, pretrained_model_name_or_path, subfolder, torch_dtype, **kwargs)

    def set_pca(self, pca_path=None, device="cuda"):
        """resampler.py:199-207 (the PCA object is a pickled top-level `pca.PCA` module)."""
        if pca_path is None:
            self.pca = None
        else:
            self.pca = torch.load(pca_path, weights_only=False).to(device)

    @staticmethod
    def _plain_ln(x2d, out2d, norm: nn.LayerNorm, ws: _Workspace):
        rows = x2d.shape[0]
        rm = E.make_rowmap(rows, 0, 0, 1, 1)
        z = E.make_modvec(ws.zeros, None, None)
        E.ln_modulate(x2d, out2d, 1, rm, norm.weight, norm.bias, None, None, norm.eps, z, z)

    def forward(self, x, image_rotary_emb=None, sampling_rotary_emb=None):
        if not x.is_cuda:
            raise E.TokensGenError("Resampler (tokensgen_b200) runs on CUDA only (no CPU fallback)")
        if self.dtype != torch.bfloat16:
            raise E.TokensGenError("tokensgen_b200 computes in bf16: call resampler.to(torch.bfloat16)")
        cfg = self.config
        dev = x.device
        B, f, n, _ = x.shape
        n_x, l, dim, H = f * n, self.latents.shape[1], cfg.dim, cfg.heads
        ff_dim = self.layers[0][1].net[0].proj.out_features
        key = (B, n_x, str(dev))
        if key not in self._tg_ws:
            self._tg_ws.clear()
            self._tg_ws[key] = _Workspace(B, n_x, l, dim, H, ff_dim, dev)
        ws = self._tg_ws[key]
        img = _rope_pair(image_rotary_emb, dev)
        smp = _rope_pair(sampling_rotary_emb, dev)
        ones = E.make_modvec(ws.ones.expand(B, dim), None, None)
        rm_lat = E.make_rowmap(l, 0, 0, 1, 1)            # gate/residual GEMMs over the latent rows only
        rm_kv = E.make_rowmap(0, n_x, l, n_x, 1)         # [x | latents]: "video" rows take the image RoPE, "vip" rows the sampling one
        rm_q = E.make_rowmap(0, 0, l, 1, 1)

        E.gemm_bias_act(x.to(torch.bfloat16).reshape(B * n_x, -1).contiguous(), self.proj_in.weight, self.proj_in.bias,
                        ws.xp.view(B * n_x, dim))
        ws.lat.copy_(self.latents.detach().expand(B, -1, -1))
        for attn, ff in self.layers:
            for b in range(B):
                self._plain_ln(ws.xp[b], ws.cat[b, :n_x], attn.norm1, ws)
                self._plain_ln(ws.lat[b], ws.cat[b, n_x:], attn.norm2, ws)
            pk, pv, pq = E.QkvProj(), E.QkvProj(), E.QkvProj()
            pk.out, pk.out_rows = ws.k.data_ptr(), n_x + l
            pv.out, pv.out_rows = ws.v.data_ptr(), n_x + l
            pq.out, pq.out_rows = ws.q.data_ptr(), l
            if attn.qk_norm:
                pk.ln_w, pk.ln_b = attn.norm_k.weight.data_ptr(), attn.norm_k.bias.data_ptr()
                pq.ln_w, pq.ln_b = attn.norm_q.weight.data_ptr(), attn.norm_q.bias.data_ptr()
            if img is not None:
                pk.cos_video, pk.sin_video = img[0].data_ptr(), img[1].data_ptr()
            if smp is not None:
                pk.cos_vip, pk.sin_vip = smp[0].data_ptr(), smp[1].data_ptr()
                pq.cos_vip, pq.sin_vip = smp[0].data_ptr(), smp[1].data_ptr()
            E.qkv_rope_gemm(ws.cat.view(B * (n_x + l), dim), attn.to_kv.weight, None, B, H, rm_kv, [pk, pv], 1e-6)
            lat_n = ws.cat[:, n_x:].contiguous().view(B * l, dim) if B > 1 else ws.cat[0, n_x:]
            E.qkv_rope_gemm(lat_n, attn.to_q.weight, None, B, H, rm_q, [pq], 1e-6)
            E.attn_fwd(ws.q, ws.k, ws.v, ws.att, softmax_scale=attn.scale)
            E.gemm_gate_residual(ws.att.view(B * l, H * 64), attn.to_out.weight, None, ws.lat.view(B * l, dim), B, rm_lat, ones)
            E.gemm_bias_act(ws.lat.view(B * l, dim), ff.net[0].proj.weight, ff.net[0].proj.bias, ws.ff, act=E.ACT_GELU_TANH)
            E.gemm_gate_residual(ws.ff, ff.net[2].weight, ff.net[2].bias, ws.lat.view(B * l, dim), B, rm_lat, ones)
        out = E.gemm_bias_act(ws.lat.view(B * l, dim), self.proj_out.weight, self.proj_out.bias)
        if out.shape[1] == dim:
            self._plain_ln(out, out, self.norm_out, ws)
        else:
            zo = torch.zeros(1, out.shape[1], device=dev, dtype=torch.bfloat16)
            rm = E.make_rowmap(out.shape[0], 0, 0, 1, 1)
            z = E.make_modvec(zo, None, None)
            E.ln_modulate(out, out, 1, rm, self.norm_out.weight, self.norm_out.bias, None, None, self.norm_out.eps, z, z)
        latents = out.view(B, l, -1)
        if self.pca is not None:
            flat = latents.reshape(B * l, -1).to(self.pca.components_.dtype)
            t = self.pca.transform(flat)
            t[:, 16:] = 0.0
            latents = self.pca.inverse_transform(t).reshape(B, l, -1).to(torch.bfloat16)
        return latents.reshape(B, self.num_temporal_queries, self.num_height_queries, self.num_width_queries, -1) \
                      .permute(0, 1, 4, 2, 3)
