"""ORACLE (test infrastructure, never shipped, never on the product path).

CPU restatement of the CogVideoX 3D causal VAE (encode / decode, frame batching with the causal conv cache, tiling and
blending) as plain functions over a state dict with the reference's key layout.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import this module.

Follows longvgen/models/autoencoder_kl_cogvideox.py (line numbers cited per function).  Three classes it uses live in
diffusers 0.31.0.dev0 (pip dependency, environment.yml:58, absent offline): CogVideoXDownsample3D, CogVideoXUpsample3D
and DiagonalGaussianDistribution; they are restated from their published semantics — that part
of the parity is UNPINNED by any reference-owned test or source in /root/reference.

Parity pin: tests/test_vae_cpu.py checks every function here against tests/golden/vae_tiny.pt, produced by
oracle/make_goldens.py running the UNMODIFIED reference AutoencoderKLCogVideoX (through oracle/stubs) on seeded inputs.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
Cache = Dict[str, Tensor]


@dataclass
class VaeConfig:
    """autoencoder_kl_cogvideox.py:922-954 (defaults = the CogVideoX-5b VAE; scaling_factor 0.7 comes from its config.json)."""
    in_channels: int = 3
    out_channels: int = 3
    block_out_channels: Tuple[int, ...] = (128, 256, 256, 512)
    latent_channels: int = 16
    layers_per_block: int = 3
    norm_num_groups: int = 32
    norm_eps: float = 1e-6
    temporal_compression_ratio: int = 4
    sample_height: int = 480
    sample_width: int = 720
    scaling_factor: float = 0.7
    num_latent_frames_batch_size: int = 2   # :1001
    num_sample_frames_batch_size: int = 8   # :1002
    tile_overlap_factor_height: float = 1 / 6  # :1011
    tile_overlap_factor_width: float = 1 / 5   # :1012

    @property
    def spatial_ratio(self) -> int:
        return 2 ** (len(self.block_out_channels) - 1)

    @property
    def tile_sample_min_height(self) -> int:
        return self.sample_height // 2  # :1007

    @property
    def tile_sample_min_width(self) -> int:
        return self.sample_width // 2

    @property
    def tile_latent_min_height(self) -> int:
        return int(self.tile_sample_min_height / self.spatial_ratio)  # :1009

    @property
    def tile_latent_min_width(self) -> int:
        return int(self.tile_sample_min_width / self.spatial_ratio)


def vae_shapes(cfg: VaeConfig) -> Dict[str, List[int]]:
    """State-dict key layout of AutoencoderKLCogVideoX (checked against the instantiated reference by make_goldens)."""
    s: Dict[str, List[int]] = {}
    g = cfg.latent_channels

    def conv3(name, cin, cout, k=3):
        s[name + ".conv.weight"] = [cout, cin, k, k, k]
        s[name + ".conv.bias"] = [cout]

    def resnet(name, cin, cout, spatial):
        for n, c in (("norm1", cin), ("norm2", cout)):
            if spatial:
                s[f"{name}.{n}.norm_layer.weight"] = [c]
                s[f"{name}.{n}.norm_layer.bias"] = [c]
                conv3(f"{name}.{n}.conv_y", g, c, 1)
                conv3(f"{name}.{n}.conv_b", g, c, 1)
            else:
                s[f"{name}.{n}.weight"] = [c]
                s[f"{name}.{n}.bias"] = [c]
        conv3(name + ".conv1", cin, cout)
        conv3(name + ".conv2", cout, cout)
        if cin != cout:
            s[name + ".conv_shortcut.weight"] = [cout, cin, 1, 1, 1]
            s[name + ".conv_shortcut.bias"] = [cout]

    boc = cfg.block_out_channels
    conv3("encoder.conv_in", cfg.in_channels, boc[0])
    out_c = boc[0]
    for i, c in enumerate(boc):
        in_c, out_c = out_c, c
        for j in range(cfg.layers_per_block):
            resnet(f"encoder.down_blocks.{i}.resnets.{j}", in_c if j == 0 else out_c, out_c, False)
        if i != len(boc) - 1:
            s[f"encoder.down_blocks.{i}.downsamplers.0.conv.weight"] = [out_c, out_c, 3, 3]
            s[f"encoder.down_blocks.{i}.downsamplers.0.conv.bias"] = [out_c]
    for j in range(2):
        resnet(f"encoder.mid_block.resnets.{j}", boc[-1], boc[-1], False)
    s["encoder.norm_out.weight"] = [boc[-1]]
    s["encoder.norm_out.bias"] = [boc[-1]]
    conv3("encoder.conv_out", boc[-1], 2 * g)

    rev = list(reversed(boc))
    conv3("decoder.conv_in", g, rev[0])
    for j in range(2):
        resnet(f"decoder.mid_block.resnets.{j}", rev[0], rev[0], True)
    out_c = rev[0]
    for i, c in enumerate(rev):
        in_c, out_c = out_c, c
        for j in range(cfg.layers_per_block + 1):
            resnet(f"decoder.up_blocks.{i}.resnets.{j}", in_c if j == 0 else out_c, out_c, True)
        if i != len(rev) - 1:
            s[f"decoder.up_blocks.{i}.upsamplers.0.conv.weight"] = [out_c, out_c, 3, 3]
            s[f"decoder.up_blocks.{i}.upsamplers.0.conv.bias"] = [out_c]
    s["decoder.norm_out.norm_layer.weight"] = [rev[-1]]
    s["decoder.norm_out.norm_layer.bias"] = [rev[-1]]
    conv3("decoder.norm_out.conv_y", g, rev[-1], 1)
    conv3("decoder.norm_out.conv_b", g, rev[-1], 1)
    conv3("decoder.conv_out", rev[-1], cfg.out_channels)
    return s


# --------------------------------------------------------------------------------------------------- layers
def causal_conv3d(sd, name: str, x: Tensor, cache: Cache) -> Tensor:
    """CogVideoXCausalConv3d.forward (:133-145) + fake_context_parallel_forward (:120-127): the time axis is padded in
    front with the last kt-1 input frames of the previous call (the conv cache) or, on the first call, with copies of
    the first frame; H and W are zero padded.  CogVideoXSafeConv3d's chunking (:44-64) re-attaches kt-1 frames of overlap
    to every chunk, so it equals the unchunked convolution and is not restated."""
    w, b = sd[name + ".conv.weight"].to(x.dtype), sd[name + ".conv.bias"].to(x.dtype)
    kt, kh, kw = w.shape[2:]
    if kt > 1:
        prev = cache.get(name)
        pad = prev if prev is not None else x[:, :, :1].repeat(1, 1, kt - 1, 1, 1)
        x = torch.cat([pad, x], dim=2)
        cache[name] = x[:, :, -(kt - 1):].clone()
    x = F.pad(x, (kw // 2, kw // 2, kh // 2, kh // 2))
    return F.conv3d(x, w, b)


def upsample_zq(zq: Tensor, size: Tuple[int, int, int]) -> Tensor:
    """Nearest interpolation of the latent onto the feature grid, first frame apart when T is odd and > 1 (:175-186)."""
    T = size[0]
    if T > 1 and T % 2 == 1:
        first = F.interpolate(zq[:, :, :1], size=(1,) + tuple(size[1:]))
        rest = F.interpolate(zq[:, :, 1:], size=(T - 1,) + tuple(size[1:]))
        return torch.cat([first, rest], dim=2)
    return F.interpolate(zq, size=tuple(size))


def spatial_norm(sd, name: str, f: Tensor, zq: Tensor, groups: int, cache: Cache) -> Tensor:
    """CogVideoXSpatialNorm3D.forward (:175-188): GroupNorm(eps 1e-6)(f) * conv_y(zq') + conv_b(zq'), zq' = upsampled zq."""
    z = upsample_zq(zq, f.shape[-3:])
    nf = F.group_norm(f, groups, sd[name + ".norm_layer.weight"].to(f.dtype), sd[name + ".norm_layer.bias"].to(f.dtype), 1e-6)
    return nf * causal_conv3d(sd, name + ".conv_y", z, cache) + causal_conv3d(sd, name + ".conv_b", z, cache)


def resnet_block(sd, name: str, x: Tensor, zq: Optional[Tensor], cfg: VaeConfig, cache: Cache) -> Tensor:
    """CogVideoXResnetBlock3D.forward (:276-309) without time embedding (temb_channels = 0 in both coders, :683,:812)."""
    g, eps = cfg.norm_num_groups, cfg.norm_eps

    def norm(n, h):
        if zq is not None:
            return spatial_norm(sd, f"{name}.{n}", h, zq, g, cache)
        return F.group_norm(h, g, sd[f"{name}.{n}.weight"].to(h.dtype), sd[f"{name}.{n}.bias"].to(h.dtype), eps)

    h = causal_conv3d(sd, name + ".conv1", F.silu(norm("norm1", x)), cache)
    h = causal_conv3d(sd, name + ".conv2", F.silu(norm("norm2", h)), cache)
    if name + ".conv_shortcut.weight" in sd:
        x = F.conv3d(x, sd[name + ".conv_shortcut.weight"].to(x.dtype), sd[name + ".conv_shortcut.bias"].to(x.dtype))
    return h + x


def downsample3d(sd, name: str, x: Tensor, compress_time: bool) -> Tensor:
    """diffusers CogVideoXDownsample3D (restated): optional avg-pool over frame pairs (first frame kept
    apart when T is odd), then F.pad(0,1,0,1) and a per-frame stride-2 3x3 convolution."""
    if compress_time:
        b, c, t, h, w = x.shape
        y = x.permute(0, 3, 4, 1, 2).reshape(b * h * w, c, t)
        if t % 2 == 1:
            first, rest = y[..., 0], y[..., 1:]
            if rest.shape[-1] > 0:
                rest = F.avg_pool1d(rest, kernel_size=2, stride=2)
            y = torch.cat([first[..., None], rest], dim=-1)
        else:
            y = F.avg_pool1d(y, kernel_size=2, stride=2)
        x = y.reshape(b, h, w, c, y.shape[-1]).permute(0, 3, 4, 1, 2)
    x = F.pad(x, (0, 1, 0, 1))
    b, c, t, h, w = x.shape
    y = F.conv2d(x.permute(0, 2, 1, 3, 4).reshape(b * t, c, h, w), sd[name + ".conv.weight"].to(x.dtype),
                 sd[name + ".conv.bias"].to(x.dtype), stride=2)
    return y.reshape(b, t, *y.shape[1:]).permute(0, 2, 1, 3, 4)


def upsample3d(sd, name: str, x: Tensor, compress_time: bool) -> Tensor:
    """diffusers CogVideoXUpsample3D (restated): nearest x2 in H, W (and in T when compress_time: first frame only in H, W
    when T is odd and > 1), then a per-frame 3x3 convolution with padding 1."""
    if compress_time:
        t = x.shape[2]
        if t > 1 and t % 2 == 1:
            first = F.interpolate(x[:, :, 0], scale_factor=2.0)
            rest = F.interpolate(x[:, :, 1:], scale_factor=2.0)
            x = torch.cat([first[:, :, None], rest], dim=2)
        elif t > 1:
            x = F.interpolate(x, scale_factor=2.0)
        else:
            x = F.interpolate(x.squeeze(2), scale_factor=2.0)[:, :, None]
    else:
        b, c, t, h, w = x.shape
        y = F.interpolate(x.permute(0, 2, 1, 3, 4).reshape(b * t, c, h, w), scale_factor=2.0)
        x = y.reshape(b, t, c, *y.shape[2:]).permute(0, 2, 1, 3, 4)
    b, c, t, h, w = x.shape
    y = F.conv2d(x.permute(0, 2, 1, 3, 4).reshape(b * t, c, h, w), sd[name + ".conv.weight"].to(x.dtype),
                 sd[name + ".conv.bias"].to(x.dtype), padding=1)
    return y.reshape(b, t, *y.shape[1:]).permute(0, 2, 1, 3, 4)


# --------------------------------------------------------------------------------------------------- coders
def encoder_forward(sd, cfg: VaeConfig, x: Tensor, cache: Cache) -> Tensor:
    """CogVideoXEncoder3D.forward (:713-745)."""
    n_time = int(math.log2(cfg.temporal_compression_ratio))
    h = causal_conv3d(sd, "encoder.conv_in", x, cache)
    nb = len(cfg.block_out_channels)
    for i in range(nb):
        for j in range(cfg.layers_per_block):
            h = resnet_block(sd, f"encoder.down_blocks.{i}.resnets.{j}", h, None, cfg, cache)
        if i != nb - 1:
            h = downsample3d(sd, f"encoder.down_blocks.{i}.downsamplers.0", h, i < n_time)
    for j in range(2):
        h = resnet_block(sd, f"encoder.mid_block.resnets.{j}", h, None, cfg, cache)
    h = F.group_norm(h, cfg.norm_num_groups, sd["encoder.norm_out.weight"].to(h.dtype), sd["encoder.norm_out.bias"].to(h.dtype), 1e-6)
    return causal_conv3d(sd, "encoder.conv_out", F.silu(h), cache)


def decoder_forward(sd, cfg: VaeConfig, z: Tensor, cache: Cache) -> Tensor:
    """CogVideoXDecoder3D.forward (:847-883): every norm is spatially conditioned on the latent itself."""
    n_time = int(math.log2(cfg.temporal_compression_ratio))
    h = causal_conv3d(sd, "decoder.conv_in", z, cache)
    for j in range(2):
        h = resnet_block(sd, f"decoder.mid_block.resnets.{j}", h, z, cfg, cache)
    nb = len(cfg.block_out_channels)
    for i in range(nb):
        for j in range(cfg.layers_per_block + 1):
            h = resnet_block(sd, f"decoder.up_blocks.{i}.resnets.{j}", h, z, cfg, cache)
        if i != nb - 1:
            h = upsample3d(sd, f"decoder.up_blocks.{i}.upsamplers.0", h, i < n_time)
    h = spatial_norm(sd, "decoder.norm_out", h, z, cfg.norm_num_groups, cache)
    return causal_conv3d(sd, "decoder.conv_out", F.silu(h), cache)


def frame_batches(num_frames: int, batch: int, single_ok: bool) -> List[Tuple[int, int]]:
    """The (start, end) frame ranges of _encode (:1091-1099, batch 8, `num_frames > 1 else 1`) and _decode (:1140-1148,
    batch 2): the remainder goes into the FIRST batch."""
    n = num_frames // batch if (num_frames > 1 or not single_ok) else 1
    rem = num_frames % batch
    return [(batch * i + (0 if i == 0 else rem), min(batch * (i + 1) + rem, num_frames)) for i in range(n)]  # slices clip


def encode_raw(sd, cfg: VaeConfig, x: Tensor) -> Tensor:
    """_encode without tiling (:1085-1107): moments [B, 2*latent, T', h, w]; conv cache carried across frame batches."""
    cache: Cache = {}
    outs = [encoder_forward(sd, cfg, x[:, :, a:b], cache) for a, b in frame_batches(x.shape[2], cfg.num_sample_frames_batch_size, True)]
    return torch.cat(outs, dim=2)


def decode_raw(sd, cfg: VaeConfig, z: Tensor) -> Tensor:
    """_decode without tiling (:1134-1163)."""
    cache: Cache = {}
    outs = [decoder_forward(sd, cfg, z[:, :, a:b], cache) for a, b in frame_batches(z.shape[2], cfg.num_latent_frames_batch_size, False)]
    return torch.cat(outs, dim=2)


def posterior_sample(moments: Tensor, eps: Tensor) -> Tensor:
    """diffusers DiagonalGaussianDistribution.sample (restated): mean + exp(0.5 * clamp(logvar, -30, 20)) * eps."""
    mean, logvar = torch.chunk(moments, 2, dim=1)
    return mean + torch.exp(0.5 * torch.clamp(logvar, -30.0, 20.0)) * eps


# --------------------------------------------------------------------------------------------------- tiling
def blend_v(a: Tensor, b: Tensor, extent: int) -> Tensor:
    """:1190-1196 (in place on b, like the reference)."""
    extent = min(a.shape[3], b.shape[3], extent)
    for y in range(extent):
        b[:, :, :, y, :] = a[:, :, :, -extent + y, :] * (1 - y / extent) + b[:, :, :, y, :] * (y / extent)
    return b


def blend_h(a: Tensor, b: Tensor, extent: int) -> Tensor:
    """:1198-1204."""
    extent = min(a.shape[4], b.shape[4], extent)
    for x in range(extent):
        b[:, :, :, :, x] = a[:, :, :, :, -extent + x] * (1 - x / extent) + b[:, :, :, :, x] * (x / extent)
    return b


def _assemble(rows, blend_h_ext, blend_v_ext, lim_h, lim_w):
    out_rows = []
    for i, row in enumerate(rows):
        out = []
        for j, tile in enumerate(row):
            if i > 0:
                tile = blend_v(rows[i - 1][j], tile, blend_v_ext)
            if j > 0:
                tile = blend_h(row[j - 1], tile, blend_h_ext)
            out.append(tile[:, :, :, :lim_h, :lim_w])
        out_rows.append(torch.cat(out, dim=4))
    return torch.cat(out_rows, dim=3)


def tiled_decode(sd, cfg: VaeConfig, z: Tensor, chunked: bool = True) -> Tensor:
    """tiled_decode (:1278-1359).  `chunked` selects the in-tree variant (13-frame chunk loop, :1317-1337); the diffusers
    runtime class uses plain frame batching — the two agree when the input holds 13 frames."""
    lh, lw = cfg.tile_latent_min_height, cfg.tile_latent_min_width
    sh, sw = cfg.tile_sample_min_height, cfg.tile_sample_min_width
    ov_h, ov_w = int(lh * (1 - cfg.tile_overlap_factor_height)), int(lw * (1 - cfg.tile_overlap_factor_width))
    be_h, be_w = int(sh * cfg.tile_overlap_factor_height), int(sw * cfg.tile_overlap_factor_width)
    lim_h, lim_w = sh - be_h, sw - be_w
    fb = cfg.num_latent_frames_batch_size
    T = z.shape[2]
    if chunked:
        ranges = [(k1 * 13 + a, k1 * 13 + b) for k1 in range(T // 13) for a, b in frame_batches(13, fb, False)]
    else:
        ranges = frame_batches(T, fb, False)
    rows = []
    for i in range(0, z.shape[3], ov_h):
        row = []
        for j in range(0, z.shape[4], ov_w):
            cache: Cache = {}
            row.append(torch.cat([decoder_forward(sd, cfg, z[:, :, a:b, i:i + lh, j:j + lw], cache) for a, b in ranges], dim=2))
        rows.append(row)
    return _assemble(rows, be_w, be_h, lim_h, lim_w)


def tiled_encode(sd, cfg: VaeConfig, x: Tensor) -> Tensor:
    """tiled_encode (:1206-1276)."""
    lh, lw = cfg.tile_latent_min_height, cfg.tile_latent_min_width
    sh, sw = cfg.tile_sample_min_height, cfg.tile_sample_min_width
    ov_h, ov_w = int(sh * (1 - cfg.tile_overlap_factor_height)), int(sw * (1 - cfg.tile_overlap_factor_width))
    be_h, be_w = int(lh * cfg.tile_overlap_factor_height), int(lw * cfg.tile_overlap_factor_width)
    lim_h, lim_w = lh - be_h, lw - be_w
    ranges = frame_batches(x.shape[2], cfg.num_sample_frames_batch_size, True)
    rows = []
    for i in range(0, x.shape[3], ov_h):
        row = []
        for j in range(0, x.shape[4], ov_w):
            cache: Cache = {}
            row.append(torch.cat([encoder_forward(sd, cfg, x[:, :, a:b, i:i + sh, j:j + sw], cache) for a, b in ranges], dim=2))
        rows.append(row)
    return _assemble(rows, be_w, be_h, lim_h, lim_w)


def encode(sd, cfg: VaeConfig, x: Tensor, tiling: bool = False) -> Tensor:
    """encode/_encode (:1085-1132): moments; tiling only when the sample exceeds the tile size (:1088)."""
    if tiling and (x.shape[4] > cfg.tile_sample_min_width or x.shape[3] > cfg.tile_sample_min_height):
        return tiled_encode(sd, cfg, x)
    return encode_raw(sd, cfg, x)


def decode(sd, cfg: VaeConfig, z: Tensor, tiling: bool = False, chunked: bool = True) -> Tensor:
    """decode/_decode (:1134-1188)."""
    if tiling and (z.shape[4] > cfg.tile_latent_min_width or z.shape[3] > cfg.tile_latent_min_height):
        return tiled_decode(sd, cfg, z, chunked)
    return decode_raw(sd, cfg, z)
