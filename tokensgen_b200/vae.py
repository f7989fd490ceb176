"""Host-side mirror of the CogVideoX 3D causal VAE, executing on the C-ABI CUDA library.

Mirrors (same class names, constructor arguments, state-dict keys, method names and return wrappers):
  longvgen/models/autoencoder_kl_cogvideox.py : CogVideoXSafeConv3d (:38-64), CogVideoXCausalConv3d (:67-145),
      CogVideoXSpatialNorm3D (:148-188), CogVideoXResnetBlock3D (:191-309), CogVideoXDownBlock3D / MidBlock3D / UpBlock3D
      (:312-608), CogVideoXEncoder3D / CogVideoXDecoder3D (:611-883), AutoencoderKLCogVideoX (:886-1377: encode, decode,
      tiled_encode, tiled_decode, blend_v, blend_h, enable_tiling, enable_slicing)
  diffusers (0.31.0.dev0, not in the reference tree): CogVideoXDownsample3D, CogVideoXUpsample3D,
      DiagonalGaussianDistribution — restated semantics, see oracle/vae.py and SURVEY.md Appendix C.

nn.Module is only the parameter container (`load_state_dict` of the public CogVideoX VAE checkpoint works key for key).
Inside the coders activations are channels-last bf16 [T, H, W, C]; every arithmetic op is a tg_vae_* / tg_gemm_* call:
  * causal conv = tcgen05 implicit GEMM reading a frame buffer whose first two frames are the conv cache (no cat / pad),
  * GroupNorm / SpatialNorm + SiLU = one statistics pass + one apply pass writing straight into the next conv's buffer,
  * the two 16->C 1x1 convolutions of SpatialNorm run once per call at latent resolution (exact: they commute with nearest
    up-sampling) instead of at feature resolution.
There is no PyTorch fallback: without the native library or a CUDA device the coders raise.
"""
from __future__ import annotations

import math
from types import SimpleNamespace
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn as nn

from . import _ext as E


def _pad64(c: int) -> int:
    return (c + 63) // 64 * 64


class _PackedConv:
    """[Cout, Cin, (kt,) kh, kw] -> [Cout_pad, kt*kh*kw*Cin_pad] tap-major K-major matrix; rebuilt when the weight changes."""

    def __init__(self):
        self.key, self.w, self.b, self.b_pad = None, None, None, None

    def get(self, conv: nn.Module):
        w = conv.weight
        key = (w.data_ptr(), w._version, None if conv.bias is None else conv.bias._version)
        if key != self.key:
            wd = w.detach()
            if wd.dim() == 4:
                wd = wd.unsqueeze(2)
            cout, cin = wd.shape[:2]
            m = wd.permute(0, 2, 3, 4, 1)
            m = torch.nn.functional.pad(m, (0, _pad64(cin) - cin))
            m = m.reshape(cout, -1)
            rows = _pad64(cout)
            m = torch.nn.functional.pad(m, (0, 0, 0, rows - cout))
            self.w = m.to(torch.bfloat16).contiguous()
            self.b = None if conv.bias is None else conv.bias.detach().to(torch.bfloat16).contiguous()
            self.b_pad = None if self.b is None else torch.nn.functional.pad(self.b, (0, rows - cout)).contiguous()
            self.key = key
        return self.w, self.b

    def linear(self, conv: nn.Module, rows: torch.Tensor) -> torch.Tensor:
        """A 1x1(x1) convolution as a GEMM over pixels: rows [P, Cin_pad] -> [P, Cout] (contiguous)."""
        w, _ = self.get(conv)
        cout = conv.weight.shape[0]
        out = E.gemm_bias_act(rows, w, self.b_pad)
        return out if out.shape[1] == cout else out[:, :cout].contiguous()


# =================================================================================================== parameter containers
class CogVideoXSafeConv3d(nn.Conv3d):
    pass


class CogVideoXCausalConv3d(nn.Module):
    def __init__(self, in_channels: int, out_channels: int, kernel_size, stride: int = 1, dilation: int = 1, pad_mode: str = "constant"):
        super().__init__()
        if isinstance(kernel_size, int):
            kernel_size = (kernel_size,) * 3
        if stride != 1 or dilation != 1:
            raise NotImplementedError("the CogVideoX VAE only uses stride 1 / dilation 1 causal convolutions")
        self.time_kernel_size = kernel_size[0]
        self.conv = CogVideoXSafeConv3d(in_channels, out_channels, kernel_size)
        self.conv_cache = None  # [kt-1, H, W, Cin_pad] channels-last: the last input frames of the previous call
        self._pack = _PackedConv()

    def _clear_fake_context_parallel_cache(self):
        self.conv_cache = None


class CogVideoXSpatialNorm3D(nn.Module):
    def __init__(self, f_channels: int, zq_channels: int, groups: int = 32):
        super().__init__()
        self.norm_layer = nn.GroupNorm(num_channels=f_channels, num_groups=groups, eps=1e-6, affine=True)
        self.conv_y = CogVideoXCausalConv3d(zq_channels, f_channels, kernel_size=1, stride=1)
        self.conv_b = CogVideoXCausalConv3d(zq_channels, f_channels, kernel_size=1, stride=1)


class CogVideoXResnetBlock3D(nn.Module):
    def __init__(self, in_channels: int, out_channels: Optional[int] = None, dropout: float = 0.0, temb_channels: int = 512,
                 groups: int = 32, eps: float = 1e-6, non_linearity: str = "swish", conv_shortcut: bool = False,
                 spatial_norm_dim: Optional[int] = None, pad_mode: str = "first"):
        super().__init__()
        out_channels = out_channels or in_channels
        if temb_channels > 0 or conv_shortcut or non_linearity not in ("swish", "silu"):
            raise NotImplementedError("VAE resnets use no time embedding, a 1x1 shortcut and SiLU")
        self.in_channels, self.out_channels = in_channels, out_channels
        if spatial_norm_dim is None:
            self.norm1 = nn.GroupNorm(num_channels=in_channels, num_groups=groups, eps=eps)
            self.norm2 = nn.GroupNorm(num_channels=out_channels, num_groups=groups, eps=eps)
        else:
            self.norm1 = CogVideoXSpatialNorm3D(in_channels, spatial_norm_dim, groups)
            self.norm2 = CogVideoXSpatialNorm3D(out_channels, spatial_norm_dim, groups)
        self.conv1 = CogVideoXCausalConv3d(in_channels, out_channels, kernel_size=3, pad_mode=pad_mode)
        self.conv2 = CogVideoXCausalConv3d(out_channels, out_channels, kernel_size=3, pad_mode=pad_mode)
        if in_channels != out_channels:
            self.conv_shortcut = CogVideoXSafeConv3d(in_channels, out_channels, kernel_size=1, stride=1, padding=0)
            self._sc_pack = _PackedConv()


class CogVideoXDownsample3D(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size=3, stride=2, padding=0, compress_time=False):
        super().__init__()
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size=kernel_size, stride=stride, padding=padding)
        self.compress_time = compress_time
        self._pack = _PackedConv()


class CogVideoXUpsample3D(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, padding=1, compress_time=False):
        super().__init__()
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size=kernel_size, stride=stride, padding=padding)
        self.compress_time = compress_time
        self._pack = _PackedConv()


class CogVideoXDownBlock3D(nn.Module):
    def __init__(self, in_channels, out_channels, temb_channels, dropout=0.0, num_layers=1, resnet_eps=1e-6, resnet_act_fn="swish",
                 resnet_groups=32, add_downsample=True, downsample_padding=0, compress_time=False, pad_mode="first"):
        super().__init__()
        self.resnets = nn.ModuleList([
            CogVideoXResnetBlock3D(in_channels if i == 0 else out_channels, out_channels, dropout, temb_channels, resnet_groups,
                                   resnet_eps, resnet_act_fn, pad_mode=pad_mode) for i in range(num_layers)])
        self.downsamplers = None
        if add_downsample:
            self.downsamplers = nn.ModuleList([CogVideoXDownsample3D(out_channels, out_channels, padding=downsample_padding,
                                                                     compress_time=compress_time)])


class CogVideoXMidBlock3D(nn.Module):
    def __init__(self, in_channels, temb_channels, dropout=0.0, num_layers=1, resnet_eps=1e-6, resnet_act_fn="swish",
                 resnet_groups=32, spatial_norm_dim=None, pad_mode="first"):
        super().__init__()
        self.resnets = nn.ModuleList([
            CogVideoXResnetBlock3D(in_channels, in_channels, dropout, temb_channels, resnet_groups, resnet_eps, resnet_act_fn,
                                   spatial_norm_dim=spatial_norm_dim, pad_mode=pad_mode) for _ in range(num_layers)])


class CogVideoXUpBlock3D(nn.Module):
    def __init__(self, in_channels, out_channels, temb_channels, dropout=0.0, num_layers=1, resnet_eps=1e-6, resnet_act_fn="swish",
                 resnet_groups=32, spatial_norm_dim=16, add_upsample=True, upsample_padding=1, compress_time=False, pad_mode="first"):
        super().__init__()
        self.resnets = nn.ModuleList([
            CogVideoXResnetBlock3D(in_channels if i == 0 else out_channels, out_channels, dropout, temb_channels, resnet_groups,
                                   resnet_eps, resnet_act_fn, spatial_norm_dim=spatial_norm_dim, pad_mode=pad_mode)
            for i in range(num_layers)])
        self.upsamplers = None
        if add_upsample:
            self.upsamplers = nn.ModuleList([CogVideoXUpsample3D(out_channels, out_channels, padding=upsample_padding,
                                                                 compress_time=compress_time)])


class CogVideoXEncoder3D(nn.Module):
    def __init__(self, in_channels=3, out_channels=16, down_block_types=("CogVideoXDownBlock3D",) * 4,
                 block_out_channels=(128, 256, 256, 512), layers_per_block=3, act_fn="silu", norm_eps=1e-6, norm_num_groups=32,
                 dropout=0.0, pad_mode="first", temporal_compression_ratio=4):
        super().__init__()
        n_time = int(math.log2(temporal_compression_ratio))
        self.conv_in = CogVideoXCausalConv3d(in_channels, block_out_channels[0], kernel_size=3, pad_mode=pad_mode)
        self.down_blocks = nn.ModuleList([])
        out_c = block_out_channels[0]
        for i, c in enumerate(block_out_channels):
            in_c, out_c = out_c, c
            self.down_blocks.append(CogVideoXDownBlock3D(in_c, out_c, 0, dropout, layers_per_block, norm_eps, act_fn, norm_num_groups,
                                                         add_downsample=i != len(block_out_channels) - 1, compress_time=i < n_time))
        self.mid_block = CogVideoXMidBlock3D(block_out_channels[-1], 0, dropout, 2, norm_eps, act_fn, norm_num_groups, pad_mode=pad_mode)
        self.norm_out = nn.GroupNorm(norm_num_groups, block_out_channels[-1], eps=1e-6)
        self.conv_act = nn.SiLU()
        self.conv_out = CogVideoXCausalConv3d(block_out_channels[-1], 2 * out_channels, kernel_size=3, pad_mode=pad_mode)


class CogVideoXDecoder3D(nn.Module):
    def __init__(self, in_channels=16, out_channels=3, up_block_types=("CogVideoXUpBlock3D",) * 4,
                 block_out_channels=(128, 256, 256, 512), layers_per_block=3, act_fn="silu", norm_eps=1e-6, norm_num_groups=32,
                 dropout=0.0, pad_mode="first", temporal_compression_ratio=4):
        super().__init__()
        rev = list(reversed(block_out_channels))
        n_time = int(math.log2(temporal_compression_ratio))
        self.conv_in = CogVideoXCausalConv3d(in_channels, rev[0], kernel_size=3, pad_mode=pad_mode)
        self.mid_block = CogVideoXMidBlock3D(rev[0], 0, num_layers=2, resnet_eps=norm_eps, resnet_act_fn=act_fn,
                                             resnet_groups=norm_num_groups, spatial_norm_dim=in_channels, pad_mode=pad_mode)
        self.up_blocks = nn.ModuleList([])
        out_c = rev[0]
        for i, c in enumerate(rev):
            in_c, out_c = out_c, c
            self.up_blocks.append(CogVideoXUpBlock3D(in_c, out_c, 0, dropout, layers_per_block + 1, norm_eps, act_fn, norm_num_groups,
                                                     spatial_norm_dim=in_channels, add_upsample=i != len(rev) - 1,
                                                     compress_time=i < n_time, pad_mode=pad_mode))
        self.norm_out = CogVideoXSpatialNorm3D(rev[-1], in_channels, groups=norm_num_groups)
        self.conv_act = nn.SiLU()
        self.conv_out = CogVideoXCausalConv3d(rev[-1], out_channels, kernel_size=3, pad_mode=pad_mode)


# =================================================================================================== engine (channels-last)
class _Act:
    """A channels-last activation [T, H, W, C] together with the GroupNorm sums of it that the producing convolution's epilogue
    accumulated (`sums` for `groups` groups; None when the producer could not, e.g. narrow widths) — so the consumer norm does
    not read the tensor once more just for its statistics (tg_vae_group_stats)."""
    __slots__ = ("x", "sums", "groups")

    def __init__(self, x, sums=None, groups=0):
        self.x, self.sums, self.groups = x, sums, groups


def _stats_for(cout: int, groups: int, device):
    if groups and E.conv_stats_supported(cout, groups):
        return torch.zeros(2 * groups, device=device, dtype=torch.float64)
    return None


def _causal_conv(mod: CogVideoXCausalConv3d, buf: torch.Tensor, T: int, cout: int, residual=None, planes_out=None, plane_stride=0,
                 stats=None, stat_groups=0):
    """buf: [kt-1+T, H, W, Cin_pad]; frames [kt-1, kt-1+T) already hold this call's input.  Fills the causal frames from the conv
    cache (or with copies of the first frame, autoencoder_kl_cogvideox.py:120-127), convolves, saves the new cache (:139)."""
    kt = mod.time_kernel_size
    w, b = mod._pack.get(mod.conv)
    _, H, W, _ = buf.shape
    if kt > 1:
        if mod.conv_cache is not None:
            buf[:kt - 1].copy_(mod.conv_cache)
        else:
            for i in range(kt - 1):
                buf[i].copy_(buf[kt - 1])
    y = E.vae_conv(buf, w, b, cout, kt, mod.conv.kernel_size[1], mod.conv.kernel_size[2], T, H, W,
                   pad_h0=mod.conv.kernel_size[1] // 2, pad_w0=mod.conv.kernel_size[2] // 2, residual=residual,
                   planes_out=planes_out, plane_stride=plane_stride, stats=stats, stat_groups=stat_groups)
    if kt > 1:
        mod.conv_cache = buf[T:T + kt - 1].clone()
    return y


class _ZqTables:
    """conv_y(zq) / conv_b(zq) of EVERY SpatialNorm of the decoder at latent resolution, from ONE GEMM: the 1x1x1 convolutions
    all read the same latent, so their weights are stacked into one [sum of 2C, 64] matrix and each norm reads its own
    column slice of the [latent pixels, sum of 2C] result (tg_norm_args.ldz).  One launch per latent batch instead of ~100
    (the tiled coder calls this for 9 tiles x 6 frame batches)."""

    def __init__(self, zq_cl: torch.Tensor, norms=None):
        self.zq = zq_cl  # [Tz, hz, wz, 64-padded latent channels]
        self.slices = None
        if norms:
            owner = norms[0]
            pack = owner.__dict__.setdefault("_tg_zq_pack", {})
            key = tuple((c.conv.weight.data_ptr(), c.conv.weight._version) for n in norms for c in (n.conv_y, n.conv_b))
            if pack.get("key") != key:
                ws, bs, offs, off = [], [], {}, 0
                for n in norms:
                    for c in (n.conv_y, n.conv_b):
                        w, _ = c._pack.get(c.conv)
                        ws.append(w)
                        bs.append(c._pack.b_pad)
                        offs[id(c)] = (off, c.conv.out_channels)
                        off += w.shape[0]
                pack.update(key=key, w=torch.cat(ws).contiguous(), b=torch.cat(bs).contiguous(), offs=offs)
            Tz, hz, wz, cp = self.zq.shape
            table = E.gemm_bias_act(self.zq.view(-1, cp), pack["w"], pack["b"]).view(Tz, hz, wz, -1)
            self.slices = {k: table[..., o:o + c] for k, (o, c) in pack["offs"].items()}

    def tables(self, norm: CogVideoXSpatialNorm3D):
        if self.slices is not None:
            return [self.slices[id(norm.conv_y)], self.slices[id(norm.conv_b)]]
        Tz, hz, wz, cp = self.zq.shape
        rows = self.zq.view(-1, cp)
        return [conv._pack.linear(conv.conv, rows).view(Tz, hz, wz, -1) for conv in (norm.conv_y, norm.conv_b)]


def _groups(norm) -> int:
    return (norm.norm_layer if isinstance(norm, CogVideoXSpatialNorm3D) else norm).num_groups


def _norm_silu_into(norm, x, out: torch.Tensor, zq: Optional[_ZqTables]):
    """x: a channels-last tensor, or an _Act whose producer already accumulated this norm's group sums."""
    if isinstance(norm, CogVideoXSpatialNorm3D):
        gn = norm.norm_layer
        zy, zb = zq.tables(norm)
    else:
        gn, zy, zb = norm, None, None
    if isinstance(x, _Act):
        sums = x.sums if (x.sums is not None and x.groups == gn.num_groups) else None
        x = x.x
    else:
        sums = None
    if sums is None:
        sums = E.vae_group_stats(x, gn.num_groups)
    E.vae_norm_act(x, sums, gn.num_groups, gn.eps, gn.weight, gn.bias, out, zy, zb, silu=True)


# GroupNorm statistics in the producing convolution's epilogue (False: the separate statistics pass, tg_vae_group_stats)
import os as _os
_FUSED_STATS = _os.environ.get("TG_VAE_FUSED_STATS", "1") != "0"
# CUDA streams the independent tiles of a tiled encode / decode are spread over (1: the reference's serial tile loop)
# Capped at 4: what has been exercised at full size (with 5 streams of tensor-core kernels active — window forwards
# beside a 3-lane tiled decode — the device hung; up to 4 never did with the shipped kernels).
_TILE_STREAMS = max(1, min(4, int(_os.environ.get("TG_VAE_TILE_STREAMS", "3"))))


def _resnet(blk: CogVideoXResnetBlock3D, xa, zq: Optional[_ZqTables], next_groups: int = 0) -> "_Act":
    """CogVideoXResnetBlock3D.forward (:276-309): norm1 -> SiLU -> conv1 -> norm2 -> SiLU -> conv2 (+ shortcut(x)).
    `xa`: tensor or _Act; returns an _Act carrying the sums the NEXT norm (`next_groups` groups) needs."""
    x = xa.x if isinstance(xa, _Act) else xa
    T, H, W, cin = x.shape
    cout = blk.out_channels
    buf = torch.empty(T + 2, H, W, cin, device=x.device, dtype=torch.bfloat16)
    _norm_silu_into(blk.norm1, xa, buf[2:], zq)
    g2 = _groups(blk.norm2)
    st1 = _stats_for(cout, g2, x.device) if _FUSED_STATS else None
    h = _causal_conv(blk.conv1, buf, T, cout, stats=st1, stat_groups=g2)
    buf2 = torch.empty(T + 2, H, W, cout, device=x.device, dtype=torch.bfloat16)
    _norm_silu_into(blk.norm2, _Act(h, st1, g2), buf2[2:], zq)
    if cin != cout:
        res = blk._sc_pack.linear(blk.conv_shortcut, x.view(-1, cin)).view(T, H, W, cout)
    else:
        res = x
    st2 = _stats_for(cout, next_groups, x.device) if _FUSED_STATS else None
    y = _causal_conv(blk.conv2, buf2, T, cout, residual=res, stats=st2, stat_groups=next_groups)
    return _Act(y, st2, next_groups)


def _conv2d(mod, x: torch.Tensor, stride: int, pad0: int, next_groups: int = 0) -> "_Act":
    T, H, W, c = x.shape
    w, b = mod._pack.get(mod.conv)
    h_out, w_out = (H // 2, W // 2) if stride == 2 else (H, W)
    cout = mod.conv.out_channels
    st = _stats_for(cout, next_groups, x.device) if _FUSED_STATS else None
    y = E.vae_conv(x, w, b, cout, 1, 3, 3, T, h_out, w_out, stride=stride, pad_h0=pad0, pad_w0=pad0, stats=st,
                   stat_groups=next_groups)
    return _Act(y, st, next_groups)


def _tile_lanes(areas, n_lanes: int):
    """Tile -> stream lane.  Tile 0 — first in the issue order: its first batch packs the weights every lane reads — stays on
    lane 0, the caller's stream; the others go largest first onto the least loaded lane (ties: lowest index)."""
    load, lane = [0] * n_lanes, [0] * len(areas)
    if areas:
        load[0] = areas[0]
    for ti in sorted(range(1, len(areas)), key=lambda i: (-areas[i], i)):
        lane[ti] = min(range(n_lanes), key=lambda k: (load[k], k))
        load[lane[ti]] += areas[ti]
    return lane


class DiagonalGaussianDistribution:
    """diffusers' posterior wrapper (restated): parameters = [mean | logvar] along dim 1."""

    def __init__(self, parameters: torch.Tensor):
        self.parameters = parameters
        self.mean, self.logvar = torch.chunk(parameters, 2, dim=1)

    def sample(self, generator: Optional[torch.Generator] = None, scale: float = 1.0) -> torch.Tensor:
        p = self.parameters
        shape = (p.shape[0], p.shape[1] // 2) + tuple(p.shape[2:])
        eps = E.randn_tensor(shape, generator, p.device, p.dtype)
        out = [E.vae_posterior_sample(p[b].contiguous(), eps[b].contiguous(), scale) for b in range(p.shape[0])]
        return torch.stack(out)

    def mode(self) -> torch.Tensor:
        return self.mean


class AutoencoderKLCogVideoX(nn.Module):
    def __init__(self, in_channels=3, out_channels=3, down_block_types=("CogVideoXDownBlock3D",) * 4,
                 up_block_types=("CogVideoXUpBlock3D",) * 4, block_out_channels=(128, 256, 256, 512), latent_channels=16,
                 layers_per_block=3, act_fn="silu", norm_eps=1e-6, norm_num_groups=32, temporal_compression_ratio=4,
                 sample_height=480, sample_width=720, scaling_factor=1.15258426, shift_factor=None, latents_mean=None,
                 latents_std=None, force_upcast=True, use_quant_conv=False, use_post_quant_conv=False):
        super().__init__()
        if use_quant_conv or use_post_quant_conv:
            raise NotImplementedError("quant_conv / post_quant_conv are not used by the CogVideoX checkpoints")
        cfg = {k: v for k, v in locals().items() if k not in ("self", "__class__")}
        self.config = SimpleNamespace(**cfg)
        if any(c % 64 for c in block_out_channels):
            raise NotImplementedError("the conv kernels need block_out_channels that are multiples of 64 (CogVideoX: 128/256/256/512)")
        self.encoder = CogVideoXEncoder3D(in_channels, latent_channels, down_block_types, block_out_channels, layers_per_block,
                                          act_fn, norm_eps, norm_num_groups, temporal_compression_ratio=temporal_compression_ratio)
        self.decoder = CogVideoXDecoder3D(latent_channels, out_channels, up_block_types, block_out_channels, layers_per_block,
                                          act_fn, norm_eps, norm_num_groups, temporal_compression_ratio=temporal_compression_ratio)
        self.quant_conv = self.post_quant_conv = None
        self.use_slicing = self.use_tiling = False
        self.num_latent_frames_batch_size = 2
        self.num_sample_frames_batch_size = 8
        self.tile_sample_min_height = sample_height // 2
        self.tile_sample_min_width = sample_width // 2
        ratio = 2 ** (len(block_out_channels) - 1)
        self.tile_latent_min_height = int(self.tile_sample_min_height / ratio)
        self.tile_latent_min_width = int(self.tile_sample_min_width / ratio)
        self.tile_overlap_factor_height = 1 / 6
        self.tile_overlap_factor_width = 1 / 5
        self.decode_chunk_frames = 13  # the in-tree tiled_decode's 13-frame chunk loop (:1317-1337)

    # ---- reference API
    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, subfolder=None, torch_dtype=None, **kwargs):
        """diffusers-style loader: <path>/<subfolder>/config.json + weights (the pipeline loads subfolder "vae")."""
        from .loading import build_from_pretrained
        return build_from_pretrained(cls, pretrained_model_name_or_path, subfolder, torch_dtype, **kwargs)

    @property
    def dtype(self):
        return self.decoder.conv_out.conv.weight.dtype

    @property
    def device(self):
        return self.decoder.conv_out.conv.weight.device

    def enable_tiling(self, tile_sample_min_height=None, tile_sample_min_width=None, tile_overlap_factor_height=None,
                      tile_overlap_factor_width=None) -> None:
        self.use_tiling = True
        self.tile_sample_min_height = tile_sample_min_height or self.tile_sample_min_height
        self.tile_sample_min_width = tile_sample_min_width or self.tile_sample_min_width
        ratio = 2 ** (len(self.config.block_out_channels) - 1)
        self.tile_latent_min_height = int(self.tile_sample_min_height / ratio)
        self.tile_latent_min_width = int(self.tile_sample_min_width / ratio)
        self.tile_overlap_factor_height = tile_overlap_factor_height or self.tile_overlap_factor_height
        self.tile_overlap_factor_width = tile_overlap_factor_width or self.tile_overlap_factor_width

    def disable_tiling(self) -> None:
        self.use_tiling = False

    def enable_slicing(self) -> None:
        self.use_slicing = True

    def disable_slicing(self) -> None:
        self.use_slicing = False

    def _clear_fake_context_parallel_cache(self):
        for m in self.modules():
            if isinstance(m, CogVideoXCausalConv3d):
                m._clear_fake_context_parallel_cache()

    def _check(self, t: torch.Tensor):
        if not t.is_cuda:
            raise E.TokensGenError("AutoencoderKLCogVideoX (tokensgen_b200) runs on CUDA only (no CPU fallback)")
        if self.dtype != torch.bfloat16:
            raise E.TokensGenError("tokensgen_b200 computes in bf16: call vae.to(torch.bfloat16)")

    # ---- coders on one frame batch, channels-last
    def _encoder_batch(self, x_cf: torch.Tensor, out_planes: torch.Tensor, plane_stride: int) -> int:
        """x_cf [3, T, H, W] (contiguous planes) -> moments written as channel planes; returns the number of latent frames."""
        enc = self.encoder
        C, T, H, W = x_cf.shape
        buf = torch.empty(T + 2, H, W, 64, device=x_cf.device, dtype=torch.bfloat16)
        E.vae_to_channels_last(x_cf, 64, buf[2:])
        G = _groups(enc.norm_out)      # every GroupNorm of a coder has the same group count (norm_num_groups)
        c0 = enc.conv_in.conv.out_channels
        st = _stats_for(c0, G, x_cf.device) if _FUSED_STATS else None
        h = _Act(_causal_conv(enc.conv_in, buf, T, c0, stats=st, stat_groups=G), st, G)
        for blk in enc.down_blocks:
            for i, r in enumerate(blk.resnets):
                last = i == len(blk.resnets) - 1 and blk.downsamplers is not None
                h = _resnet(r, h, None, 0 if last else G)      # a down-sampler, not a norm, reads the block's last output
            if blk.downsamplers is not None:
                d = blk.downsamplers[0]
                x = h.x
                if d.compress_time:
                    x = E.vae_avgpool_time(x)
                h = _conv2d(d, x, 2, 0, G)
        for r in enc.mid_block.resnets:
            h = _resnet(r, h, None, G)
        T2, H2, W2, c = h.x.shape
        buf = torch.empty(T2 + 2, H2, W2, c, device=h.x.device, dtype=torch.bfloat16)
        _norm_silu_into(enc.norm_out, h, buf[2:], None)
        _causal_conv(enc.conv_out, buf, T2, enc.conv_out.conv.out_channels, planes_out=out_planes, plane_stride=plane_stride)
        return T2

    def _spatial_norms(self):
        lst = self.__dict__.get("_tg_spatial_norms")
        if lst is None:
            lst = [m for m in self.decoder.modules() if isinstance(m, CogVideoXSpatialNorm3D)]
            self.__dict__["_tg_spatial_norms"] = lst
        return lst

    def _decoder_batch(self, z_cf: torch.Tensor, out_planes: torch.Tensor, plane_stride: int) -> int:
        """z_cf [16, T, h, w] -> video frames written as channel planes; returns the number of frames produced."""
        dec = self.decoder
        C, T, H, W = z_cf.shape
        buf = torch.empty(T + 2, H, W, 64, device=z_cf.device, dtype=torch.bfloat16)
        E.vae_to_channels_last(z_cf, 64, buf[2:])
        zq = _ZqTables(buf[2:], self._spatial_norms())
        G = _groups(dec.norm_out)
        c0 = dec.conv_in.conv.out_channels
        st = _stats_for(c0, G, z_cf.device) if _FUSED_STATS else None
        h = _Act(_causal_conv(dec.conv_in, buf, T, c0, stats=st, stat_groups=G), st, G)
        # buf[2:] (zq) stays alive and unmodified: the conv only rewrites the two causal frames in front of it
        for r in dec.mid_block.resnets:
            h = _resnet(r, h, zq, G)
        for blk in dec.up_blocks:
            for i, r in enumerate(blk.resnets):
                last = i == len(blk.resnets) - 1 and blk.upsamplers is not None
                h = _resnet(r, h, zq, 0 if last else G)        # an up-sampler, not a norm, reads the block's last output
            if blk.upsamplers is not None:
                u = blk.upsamplers[0]
                h = _conv2d(u, E.vae_upsample(h.x, u.compress_time), 1, 1, G)
        T2, H2, W2, c = h.x.shape
        buf2 = torch.empty(T2 + 2, H2, W2, c, device=h.x.device, dtype=torch.bfloat16)
        _norm_silu_into(dec.norm_out, h, buf2[2:], zq)
        _causal_conv(dec.conv_out, buf2, T2, dec.conv_out.conv.out_channels, planes_out=out_planes, plane_stride=plane_stride)
        return T2

    @staticmethod
    def _frame_batches(num_frames: int, batch: int, single_ok: bool) -> List[Tuple[int, int]]:
        n = num_frames // batch if (num_frames > 1 or not single_ok) else 1
        rem = num_frames % batch
        return [(batch * i + (0 if i == 0 else rem), min(batch * (i + 1) + rem, num_frames)) for i in range(n)]

    def _latent_frames(self, t: int) -> int:
        """Frames one encoder call yields: every compress_time level pools frame pairs and keeps an odd first frame."""
        for _ in range(int(math.log2(self.config.temporal_compression_ratio))):
            t = 1 + (t - 1) // 2 if t % 2 == 1 else t // 2
        return t

    def _sample_frames(self, t: int) -> int:
        """Frames one decoder call yields: every compress_time level doubles T (2T-1 when T is odd and > 1)."""
        for _ in range(int(math.log2(self.config.temporal_compression_ratio))):
            t = t if t == 1 else (2 * t - 1 if t % 2 == 1 else 2 * t)
        return t

    def _encode_one(self, x: torch.Tensor, ranges) -> torch.Tensor:
        """x [3, T, H, W] -> moments [2*latent, T', h, w] (conv cache carried across `ranges`, cleared after)."""
        C, T, H, W = x.shape
        ratio = 2 ** (len(self.config.block_out_channels) - 1)
        t_lat = sum(self._latent_frames(b - a) for a, b in ranges)
        out = torch.empty(2 * self.config.latent_channels, t_lat, H // ratio, W // ratio, device=x.device, dtype=torch.bfloat16)
        t0 = 0
        for a, b in ranges:
            t0 += self._encoder_batch(x[:, a:b].contiguous(), out[:, t0:], out.stride(0))
        self._clear_fake_context_parallel_cache()
        assert t0 == t_lat, (t0, t_lat)
        return out

    # ---- tiling (:1206-1359)
    def _assemble(self, rows, be_w, be_h, lim_h, lim_w) -> torch.Tensor:
        out_rows = []
        for i, row in enumerate(rows):
            parts = []
            for j, tile in enumerate(row):
                if i > 0:
                    E.vae_blend(rows[i - 1][j], tile, be_h, 0)
                if j > 0:
                    E.vae_blend(row[j - 1], tile, be_w, 1)
                parts.append(tile[:, :, :lim_h, :lim_w])
            out_rows.append(torch.cat(parts, dim=3))
        return torch.cat(out_rows, dim=2)

    def _causal_convs(self):
        lst = self.__dict__.get("_tg_causal_convs")
        if lst is None:
            lst = [m for m in self.modules() if isinstance(m, CogVideoXCausalConv3d)]
            self.__dict__["_tg_causal_convs"] = lst
        return lst

    def _run_tiles(self, tiles, ranges, batch_fn, alloc_fn):
        """The tiles of a tiled encode / decode are independent causal streams (the reference runs them one after the other,
        :1256-1285, :1317-1340).  Here their frame batches are issued round-robin onto `_TILE_STREAMS` CUDA streams, each tile
        with its own conv-cache state, so one tile's small launches (a 30 x 45 latent tile fills 64 of 148 SMs in the 512-channel
        blocks), launch gaps and wave tails are covered by the other tiles' kernels.  Every kernel's result is independent of
        what runs beside it: the outputs are those of the serial loop.
        tiles: input tensors [C, T, h, w]; alloc_fn(tile) -> output planes; batch_fn(tile[:, a:b], out[:, t0:], stride) -> frames."""
        convs = self._causal_convs()
        n_streams = max(1, min(_TILE_STREAMS, len(tiles)))
        main = torch.cuda.current_stream()
        if n_streams > 1:
            pool = self.__dict__.setdefault("_tg_tile_streams", {})
            key = (main.device, n_streams)
            if key not in pool:
                pool[key] = [torch.cuda.Stream(device=main.device) for _ in range(n_streams - 1)]
            streams = [main] + pool[key]
        else:
            streams = [main]
        outs = [alloc_fn(t) for t in tiles]
        state = [[None] * len(convs) for _ in tiles]
        t0 = [0] * len(tiles)
        # tiles differ in size (edge tiles are narrower / shorter): largest first onto the least loaded stream; tile 0 (a full
        # tile, first in the order) lands on the main stream
        lane = _tile_lanes([t.shape[2] * t.shape[3] for t in tiles], n_streams)
        forked = n_streams == 1
        for a, b in ranges:
            for ti, tile in enumerate(tiles):
                if not forked and ti == 1:
                    # the first batch of tile 0 (main stream) has packed the weights / zq tables every stream reads from here on
                    for st in streams[1:]:
                        st.wait_stream(main)
                    forked = True
                with torch.cuda.stream(streams[lane[ti]]):
                    for m, c in zip(convs, state[ti]):
                        m.conv_cache = c
                    t0[ti] += batch_fn(tile[:, a:b].contiguous(), outs[ti][:, t0[ti]:], outs[ti].stride(0))
                    state[ti] = [m.conv_cache for m in convs]
        for st in streams[1:]:
            main.wait_stream(st)
        self._clear_fake_context_parallel_cache()
        for o, n in zip(outs, t0):
            assert n == o.shape[1], (n, o.shape)
        return outs

    def _tiled_decode_one(self, z: torch.Tensor) -> torch.Tensor:
        lh, lw, sh, sw = self.tile_latent_min_height, self.tile_latent_min_width, self.tile_sample_min_height, self.tile_sample_min_width
        ov_h, ov_w = int(lh * (1 - self.tile_overlap_factor_height)), int(lw * (1 - self.tile_overlap_factor_width))
        be_h, be_w = int(sh * self.tile_overlap_factor_height), int(sw * self.tile_overlap_factor_width)
        T = z.shape[1]
        nf = self.decode_chunk_frames
        fb = self.num_latent_frames_batch_size
        if T % nf == 0:  # in-tree variant (:1317-1337): 13-frame chunks back to back, cache NOT cleared in between
            ranges = [(k1 * nf + a, k1 * nf + b) for k1 in range(T // nf) for a, b in self._frame_batches(nf, fb, False)]
        else:            # the diffusers runtime class: plain frame batching (identical for 13 frames)
            ranges = self._frame_batches(T, fb, False)
        ii, jj = list(range(0, z.shape[2], ov_h)), list(range(0, z.shape[3], ov_w))
        tiles = [z[:, :, i:i + lh, j:j + lw] for i in ii for j in jj]
        ratio = 2 ** (len(self.config.block_out_channels) - 1)
        n_out = sum(self._sample_frames(b - a) for a, b in ranges)
        alloc = lambda t: torch.empty(self.config.out_channels, n_out, t.shape[2] * ratio, t.shape[3] * ratio, device=z.device,
                                      dtype=torch.bfloat16)
        outs = self._run_tiles(tiles, ranges, self._decoder_batch, alloc)
        rows = [outs[r * len(jj):(r + 1) * len(jj)] for r in range(len(ii))]
        return self._assemble(rows, be_w, be_h, sh - be_h, sw - be_w)

    def _decode_stream(self, z: torch.Tensor, ranges) -> torch.Tensor:
        """One causal stream over `ranges` (cache carried through all of them, like the reference's tile loop)."""
        ratio = 2 ** (len(self.config.block_out_channels) - 1)
        n_out = sum(self._sample_frames(b - a) for a, b in ranges)
        out = torch.empty(self.config.out_channels, n_out, z.shape[2] * ratio, z.shape[3] * ratio, device=z.device, dtype=torch.bfloat16)
        t0 = 0
        for a, b in ranges:
            t0 += self._decoder_batch(z[:, a:b].contiguous(), out[:, t0:], out.stride(0))
        self._clear_fake_context_parallel_cache()
        assert t0 == n_out, (t0, n_out)
        return out

    def _tiled_encode_one(self, x: torch.Tensor) -> torch.Tensor:
        lh, lw, sh, sw = self.tile_latent_min_height, self.tile_latent_min_width, self.tile_sample_min_height, self.tile_sample_min_width
        ov_h, ov_w = int(sh * (1 - self.tile_overlap_factor_height)), int(sw * (1 - self.tile_overlap_factor_width))
        be_h, be_w = int(lh * self.tile_overlap_factor_height), int(lw * self.tile_overlap_factor_width)
        ranges = self._frame_batches(x.shape[1], self.num_sample_frames_batch_size, True)
        ii, jj = list(range(0, x.shape[2], ov_h)), list(range(0, x.shape[3], ov_w))
        tiles = [x[:, :, i:i + sh, j:j + sw] for i in ii for j in jj]
        ratio = 2 ** (len(self.config.block_out_channels) - 1)
        t_lat = sum(self._latent_frames(b - a) for a, b in ranges)
        alloc = lambda t: torch.empty(2 * self.config.latent_channels, t_lat, t.shape[2] // ratio, t.shape[3] // ratio,
                                      device=x.device, dtype=torch.bfloat16)
        outs = self._run_tiles(tiles, ranges, self._encoder_batch, alloc)
        rows = [outs[r * len(jj):(r + 1) * len(jj)] for r in range(len(ii))]
        return self._assemble(rows, be_w, be_h, lh - be_h, lw - be_w)

    # ---- public encode / decode (:1085-1188)
    def _encode(self, x: torch.Tensor) -> torch.Tensor:
        outs = []
        for xb in x:
            if self.use_tiling and (xb.shape[3] > self.tile_sample_min_width or xb.shape[2] > self.tile_sample_min_height):
                outs.append(self._tiled_encode_one(xb))
            else:
                outs.append(self._encode_one(xb, self._frame_batches(xb.shape[1], self.num_sample_frames_batch_size, True)))
        return torch.stack(outs)

    def encode(self, x: torch.Tensor, return_dict: bool = True):
        self._check(x)
        posterior = DiagonalGaussianDistribution(self._encode(x.to(torch.bfloat16)))
        if not return_dict:
            return (posterior,)
        return SimpleNamespace(latent_dist=posterior)

    def _decode(self, z: torch.Tensor):
        outs = []
        for zb in z:
            if self.use_tiling and (zb.shape[3] > self.tile_latent_min_width or zb.shape[2] > self.tile_latent_min_height):
                outs.append(self._tiled_decode_one(zb))
            else:
                outs.append(self._decode_stream(zb, self._frame_batches(zb.shape[1], self.num_latent_frames_batch_size, False)))
        return torch.stack(outs)

    def decode(self, z: torch.Tensor, return_dict: bool = True):
        self._check(z)
        dec = self._decode(z.to(torch.bfloat16))
        if not return_dict:
            return (dec,)
        return SimpleNamespace(sample=dec)

    def forward(self, sample: torch.Tensor, sample_posterior: bool = False, return_dict: bool = True,
                generator: Optional[torch.Generator] = None):
        posterior = self.encode(sample).latent_dist
        z = posterior.sample(generator=generator) if sample_posterior else posterior.mode()
        dec = self.decode(z)
        return dec if return_dict else (dec,)
