"""GPU parity of the 3D causal VAE: CUDA path (through the C-ABI) vs the oracle (oracle/vae.py, pinned to the unmodified
reference by tests/test_vae_cpu.py).

Tolerances: bf16 kernels with fp32 accumulation vs the fp32 oracle on the same bf16-rounded inputs/weights —
relative L2 <= 5e-3 per op, <= 1.5e-2 for a whole coder (~20-40 conv layers deep; the reference's own bf16 run is at ~1e-2 on
the golden case) and PSNR >= 40 dB on decoded frames (longvgen/metrics/psnr_ssim.py:50-78 on [0, 1]
images).  Resampling, layout maps and blending are bit-exact.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu

CFG = dict(block_out_channels=(64, 128, 128, 128), latent_channels=16, layers_per_block=1, norm_num_groups=32,
           sample_height=96, sample_width=80, scaling_factor=0.7)


def rel_l2(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm()).item()


def psnr(a, b):
    """calculate_psnr_pt (longvgen/metrics/psnr_ssim.py:50-78) on the post-processed video: frames mapped to [0, 1]."""
    a, b = (a.double().cpu() / 2 + 0.5).clamp(0, 1), (b.double().cpu() / 2 + 0.5).clamp(0, 1)
    return (10.0 * torch.log10(1.0 / (((a - b) ** 2).mean() + 1e-8))).item()


@pytest.fixture(scope="module")
def env():
    from oracle import vae as ov
    from oracle.synth import synth_state_dict
    from tokensgen_b200.vae import AutoencoderKLCogVideoX
    cfg = ov.VaeConfig(**CFG)
    sd = synth_state_dict(ov.vae_shapes(cfg), seed=99)
    vae = AutoencoderKLCogVideoX(block_out_channels=cfg.block_out_channels, latent_channels=16, layers_per_block=1,
                                 norm_num_groups=32, sample_height=96, sample_width=80, scaling_factor=0.7)
    vae.load_state_dict(sd, strict=True)
    vae = vae.to("cuda", torch.bfloat16).eval()
    return ov, cfg, {k: v.float() for k, v in sd.items()}, vae


def cl(x):  # [1,C,T,H,W] -> channels-last [T,H,W,C] on the GPU
    return x[0].permute(1, 2, 3, 0).contiguous().cuda()


def cf(y):  # channels-last [T,H,W,C] -> [1,C,T,H,W]
    return y.permute(3, 0, 1, 2).unsqueeze(0)


def test_causal_conv_with_cache_vs_oracle(env):
    ov, cfg, sd, vae = env
    from tokensgen_b200 import vae as V
    g = torch.Generator().manual_seed(0)
    mod = vae.decoder.up_blocks[0].resnets[0].conv1  # 128 -> 128
    name = "decoder.up_blocks.0.resnets.0.conv1"
    cache = {}
    mod._clear_fake_context_parallel_cache()
    for T, H, W in ((3, 12, 10), (2, 12, 10)):  # second call consumes the first call's cache
        x = torch.randn(1, 128, T, H, W, generator=g).bfloat16()
        ref = ov.causal_conv3d(sd, name, x.float(), cache)
        buf = torch.empty(T + 2, H, W, 128, device="cuda", dtype=torch.bfloat16)
        buf[2:] = cl(x)
        y = V._causal_conv(mod, buf, T, 128)
        assert rel_l2(cf(y), ref) < 5e-3
    mod._clear_fake_context_parallel_cache()


def test_conv_tile_edges_stride2_and_residual(env):
    ov, cfg, sd, vae = env
    from tokensgen_b200 import _ext as E
    from tokensgen_b200.vae import _PackedConv
    g = torch.Generator().manual_seed(1)
    # 3x3 per-frame conv, odd sizes that do not fill the 32x8 tiles, with a residual
    conv = torch.nn.Conv2d(64, 128, 3, padding=1)
    with torch.no_grad():
        conv.weight.copy_(torch.randn(conv.weight.shape, generator=g) / 24)
    conv = conv.to(torch.bfloat16)
    x = torch.randn(2, 64, 37, 45, generator=g).bfloat16()
    res = torch.randn(2, 37, 45, 128, generator=g).bfloat16()
    ref = torch.nn.functional.conv2d(x.float(), conv.weight.float(), conv.bias.float(), padding=1).permute(0, 2, 3, 1) + res.float()
    w, b = _PackedConv().get(conv.cuda())
    y = E.vae_conv(x.permute(0, 2, 3, 1).contiguous().cuda(), w, b, 128, 1, 3, 3, 2, 37, 45, residual=res.cuda())
    assert rel_l2(y, ref) < 5e-3
    # stride-2 down-sampling conv with the (0,1,0,1) padding of CogVideoXDownsample3D
    x = torch.randn(3, 64, 24, 40, generator=g).bfloat16()
    ref = torch.nn.functional.conv2d(torch.nn.functional.pad(x.float(), (0, 1, 0, 1)), conv.weight.float().cpu(),
                                     conv.bias.float().cpu(), stride=2).permute(0, 2, 3, 1)
    y = E.vae_conv(x.permute(0, 2, 3, 1).contiguous().cuda(), w, b, 128, 1, 3, 3, 3, 12, 20, stride=2, pad_h0=0, pad_w0=0)
    assert rel_l2(y, ref) < 5e-3


@pytest.mark.parametrize("cin,cout,T,H,W", [(128, 256, 3, 8, 32), (64, 256, 1, 19, 70), (256, 512, 2, 9, 33), (128, 128, 3, 8, 32)])
def test_conv_cta_pair_tiles(cin, cout, T, H, W):
    """The CTA-pair kernels: 256-channel pair tiles (one accumulator buffer), 128-channel ones (two), and an ODD number of
    256-pixel tiles (the last pair's second CTA runs a padding tile whose loads zero-fill and whose stores are masked)."""
    from tokensgen_b200 import _ext as E
    from tokensgen_b200.vae import _PackedConv
    g = torch.Generator().manual_seed(cin + cout + W)
    conv = torch.nn.Conv3d(cin, cout, 3)
    with torch.no_grad():
        conv.weight.copy_(torch.randn(conv.weight.shape, generator=g) / (27 * cin) ** 0.5)
        conv.bias.copy_(torch.randn(cout, generator=g))
    conv = conv.to(torch.bfloat16)
    x = torch.randn(1, cin, T + 2, H, W, generator=g).bfloat16()      # the two leading frames are the causal context
    ref = torch.nn.functional.conv3d(torch.nn.functional.pad(x.float(), (1, 1, 1, 1)), conv.weight.float(), conv.bias.float())
    w, b = _PackedConv().get(conv.cuda())
    y = E.vae_conv(x[0].permute(1, 2, 3, 0).contiguous().cuda(), w, b, cout, 3, 3, 3, T, H, W)
    torch.cuda.synchronize()
    assert rel_l2(y.permute(3, 0, 1, 2).unsqueeze(0), ref) < 5e-3


@pytest.mark.parametrize("cin,cout,T,H,W,res", [(64, 128, 2, 160, 250, False), (128, 256, 1, 150, 300, True), (128, 128, 3, 97, 129, True)])
def test_conv_tap_reuse_kernel(cin, cout, T, H, W, res):
    """conv3_kernel (CTA pairs + one halo'd activation box per (kt, kh) serving the three kw taps through swizzle-phase
    descriptors): engages on stride-1 3x3x3 layers with >= two waves of 128 x 2 pixel tiles.  Ragged widths / heights, an odd
    number of tiles, two channel tiles and the residual epilogue."""
    from tokensgen_b200 import _ext as E
    from tokensgen_b200.vae import _PackedConv
    g = torch.Generator().manual_seed(cin * 7 + W)
    conv = torch.nn.Conv3d(cin, cout, 3)
    with torch.no_grad():
        conv.weight.copy_(torch.randn(conv.weight.shape, generator=g) / (27 * cin) ** 0.5)
        conv.bias.copy_(torch.randn(cout, generator=g))
    conv = conv.to(torch.bfloat16)
    x = torch.randn(1, cin, T + 2, H, W, generator=g).bfloat16()
    r = torch.randn(T, H, W, cout, generator=g).bfloat16() if res else None
    ref = torch.nn.functional.conv3d(torch.nn.functional.pad(x.float(), (1, 1, 1, 1)), conv.weight.float(), conv.bias.float())
    ref = ref[0].permute(1, 2, 3, 0)
    if res:
        ref = ref + r.float()
    w, b = _PackedConv().get(conv.cuda())
    y = E.vae_conv(x[0].permute(1, 2, 3, 0).contiguous().cuda(), w, b, cout, 3, 3, 3, T, H, W, residual=None if r is None else r.cuda())
    torch.cuda.synchronize()
    assert rel_l2(y, ref) < 5e-3


@pytest.mark.parametrize("T,H,W", [(3, 150, 301), (8, 240, 360), (1, 20, 40)])
def test_narrow_output_conv(T, H, W):
    """The decoder's conv_out (128 -> 3, channel-plane output).  At full resolution the layer is bound by the delivery of its
    activation boxes, so it runs on the tap-reuse CTA-pair kernel with ONE 32-column tile (conv3_kernel<32>: 9 boxes per pixel
    tile instead of 27, N = 32 MMAs instead of 64 padded columns); small maps stay on the single-CTA kernel.  Against torch
    conv3d; the zero-padded weight rows must not leak into the three planes or past them."""
    from tokensgen_b200 import _ext as E
    from tokensgen_b200.vae import _PackedConv
    g = torch.Generator().manual_seed(77 + W)
    cin, cout = 128, 3
    conv = torch.nn.Conv3d(cin, cout, 3)
    with torch.no_grad():
        conv.weight.copy_(torch.randn(conv.weight.shape, generator=g) / (27 * cin) ** 0.5)
        conv.bias.copy_(torch.randn(cout, generator=g))
    conv = conv.to(torch.bfloat16).cuda()
    x = torch.randn(1, cin, T + 2, H, W, generator=g).bfloat16()
    ref = torch.nn.functional.conv3d(torch.nn.functional.pad(x.float(), (1, 1, 1, 1)).cuda(), conv.weight.float(), conv.bias.float())[0]
    xc = x[0].permute(1, 2, 3, 0).contiguous().cuda()
    w, b = _PackedConv().get(conv)
    planes = torch.full((cout, T + 1, H, W), 7.0, device="cuda", dtype=torch.bfloat16)   # one spare frame: must stay untouched
    E.vae_conv(xc, w, b, cout, 3, 3, 3, T, H, W, planes_out=planes, plane_stride=planes.stride(0))
    torch.cuda.synchronize()
    assert (planes[:, T] == 7.0).all()
    assert rel_l2(planes[:, :T], ref) < 5e-3
    # and channels-last with a row pitch (the generic epilogue path)
    y = E.vae_conv(xc, w, b, cout, 3, 3, 3, T, H, W)
    assert y.shape == (T, H, W, cout) and rel_l2(y.permute(3, 0, 1, 2), ref) < 5e-3


def test_spatial_norm_silu_vs_oracle(env):
    ov, cfg, sd, vae = env
    from tokensgen_b200 import vae as V
    g = torch.Generator().manual_seed(2)
    norm = vae.decoder.up_blocks[0].resnets[0].norm1
    name = "decoder.up_blocks.0.resnets.0.norm1"
    for (T, Tz) in ((5, 3), (4, 2), (3, 3), (1, 1)):
        f = torch.randn(1, 128, T, 8, 12, generator=g).bfloat16()
        zq = torch.randn(1, 16, Tz, 2, 3, generator=g).bfloat16()
        ref = torch.nn.functional.silu(ov.spatial_norm(sd, name, f.float(), zq.float(), 32, {}))
        zq_cl = torch.zeros(Tz, 2, 3, 64, device="cuda", dtype=torch.bfloat16)
        zq_cl[..., :16] = cl(zq)
        out = torch.empty(T, 8, 12, 128, device="cuda", dtype=torch.bfloat16)
        V._norm_silu_into(norm, cl(f), out, V._ZqTables(zq_cl))
        assert rel_l2(cf(out), ref) < 5e-3, (T, Tz)
    # plain GroupNorm + SiLU (encoder)
    gn = vae.encoder.norm_out
    f = torch.randn(1, 128, 3, 6, 5, generator=g).bfloat16() * 2 + 0.5
    ref = torch.nn.functional.silu(torch.nn.functional.group_norm(f.float(), 32, sd["encoder.norm_out.weight"], sd["encoder.norm_out.bias"], 1e-6))
    out = torch.empty(3, 6, 5, 128, device="cuda", dtype=torch.bfloat16)
    V._norm_silu_into(gn, cl(f), out, None)
    assert rel_l2(cf(out), ref) < 5e-3


def test_resampling_and_layout_bit_exact():
    from tokensgen_b200 import _ext as E
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(3)
    for T in (3, 2, 1, 5):
        x = torch.randn(1, 64, T, 5, 7, generator=g).bfloat16()
        # compress_time semantics restated from diffusers (oracle.upsample3d without the conv)
        if T > 1 and T % 2 == 1:
            ref = torch.cat([F.interpolate(x[:, :, 0], scale_factor=2.0)[:, :, None], F.interpolate(x[:, :, 1:], scale_factor=2.0)], 2)
        elif T > 1:
            ref = F.interpolate(x, scale_factor=2.0)
        else:
            ref = F.interpolate(x.squeeze(2), scale_factor=2.0)[:, :, None]
        assert torch.equal(cf(E.vae_upsample(cl(x), True)).cpu(), ref)
        ref = torch.stack([F.interpolate(x[:, :, t], scale_factor=2.0) for t in range(T)], 2)
        assert torch.equal(cf(E.vae_upsample(cl(x), False)).cpu(), ref)
    for T in (9, 8, 1, 2):
        x = torch.randn(1, 64, T, 4, 6, generator=g).bfloat16()
        y = x.permute(0, 3, 4, 1, 2).reshape(24, 64, T)
        if T % 2 == 1:
            rest = F.avg_pool1d(y[..., 1:].float(), 2, 2).bfloat16() if T > 1 else y[..., 1:]
            y = torch.cat([y[..., :1], rest], -1)
        else:
            y = F.avg_pool1d(y.float(), 2, 2).bfloat16()
        ref = y.reshape(1, 4, 6, 64, -1).permute(0, 3, 4, 1, 2)
        assert torch.equal(cf(E.vae_avgpool_time(cl(x))).cpu(), ref)
    x = torch.randn(3, 4, 6, 10, generator=g).bfloat16().cuda()
    y = E.vae_to_channels_last(x, 64)
    assert torch.equal(y[..., :3], x.permute(1, 2, 3, 0)) and not y[..., 3:].any()


def test_blend_bit_exact_vs_torch_on_gpu():
    """blend_v / blend_h written exactly like autoencoder_kl_cogvideox.py:1190-1204, evaluated by torch on this GPU."""
    from tokensgen_b200 import _ext as E
    g = torch.Generator().manual_seed(4)
    a = torch.randn(3, 5, 30, 24, generator=g).bfloat16().cuda()
    b = torch.randn(3, 5, 22, 24, generator=g).bfloat16().cuda()
    ref = b.clone()
    ext = min(a.shape[2], b.shape[2], 8)
    for y in range(ext):
        ref[:, :, y, :] = a[:, :, -ext + y, :] * (1 - y / ext) + ref[:, :, y, :] * (y / ext)
    E.vae_blend(a, b, 8, 0)
    assert torch.equal(b, ref)
    a = torch.randn(3, 5, 22, 30, generator=g).bfloat16().cuda()
    ref = b.clone()
    ext = min(a.shape[3], b.shape[3], 40)  # clamps to 24
    for x in range(ext):
        ref[:, :, :, x] = a[:, :, :, -ext + x] * (1 - x / ext) + ref[:, :, :, x] * (x / ext)
    E.vae_blend(a, b, 40, 1)
    assert torch.equal(b, ref)


def test_encode_decode_vs_oracle(env):
    ov, cfg, sd, vae = env
    g = torch.Generator().manual_seed(5)
    x = (torch.rand(1, 3, 17, 32, 40, generator=g) * 2 - 1).bfloat16()
    vae.disable_tiling()
    m = vae.encode(x.cuda()).latent_dist.parameters
    ref = ov.encode(sd, cfg, x.float())
    e = rel_l2(m, ref)
    z = torch.randn(1, 16, 5, 4, 5, generator=g).bfloat16()
    d = vae.decode(z.cuda()).sample
    refd = ov.decode(sd, cfg, z.float())
    ed = rel_l2(d, refd)
    print(f"encode rel_l2 {e:.3e}  decode rel_l2 {ed:.3e}  decode PSNR {psnr(d, refd):.1f} dB")
    assert m.shape == ref.shape and d.shape == refd.shape == (1, 3, 17, 32, 40)
    assert e < 1.5e-2 and ed < 1.5e-2 and psnr(d, refd) >= 40.0
    # encode -> sample -> decode round trip runs and the fused posterior sample matches its definition
    eps = torch.randn(16, 5, 4, 5, generator=g).bfloat16().cuda()
    from tokensgen_b200 import _ext as E
    zs = E.vae_posterior_sample(m[0].contiguous(), eps, 0.7)
    mean, logvar = m[0].float().chunk(2, dim=0)
    want = ((mean + torch.exp(0.5 * logvar.clamp(-30, 20)).bfloat16().float() * eps.float()).bfloat16().float() * 0.7).bfloat16()
    assert rel_l2(zs, want) < 1e-3


def test_tiled_decode_and_encode_vs_oracle(env):
    ov, cfg, sd, vae = env
    g = torch.Generator().manual_seed(6)
    z = torch.randn(1, 16, 13, 12, 10, generator=g).bfloat16()
    vae.enable_tiling()
    d = vae.decode(z.cuda()).sample
    ref = ov.decode(sd, cfg, z.float(), tiling=True)
    x = (torch.rand(1, 3, 9, 96, 80, generator=g) * 2 - 1).bfloat16()
    m = vae.encode(x.cuda()).latent_dist.parameters
    refm = ov.encode(sd, cfg, x.float(), tiling=True)
    vae.disable_tiling()
    print(f"tiled decode rel_l2 {rel_l2(d, ref):.3e} PSNR {psnr(d, ref):.1f} dB  tiled encode rel_l2 {rel_l2(m, refm):.3e}")
    assert d.shape == ref.shape == (1, 3, 49, 96, 80) and m.shape == refm.shape
    assert rel_l2(d, ref) < 1.5e-2 and rel_l2(m, refm) < 1.5e-2 and psnr(d, ref) >= 40.0


def test_tile_streams_do_not_change_the_tiled_coder(env):
    """The tiles of a tiled decode / encode are spread over CUDA streams (vae._run_tiles); the frames must be the ones the
    serial tile loop of the reference order (TG_VAE_TILE_STREAMS=1) produces, bit for bit."""
    ov, cfg, sd, vae = env
    from tokensgen_b200 import vae as V
    g = torch.Generator().manual_seed(16)
    z = torch.randn(1, 16, 13, 12, 10, generator=g).bfloat16().cuda()
    x = (torch.rand(1, 3, 9, 96, 80, generator=g) * 2 - 1).bfloat16().cuda()
    vae.enable_tiling()
    outs = {}
    keep = V._TILE_STREAMS
    try:
        for n in (1, 3, 4):
            V._TILE_STREAMS = n
            outs[n] = (vae.decode(z).sample.clone(), vae.encode(x).latent_dist.parameters.clone())
    finally:
        V._TILE_STREAMS = keep
        vae.disable_tiling()
    for n in (3, 4):
        assert torch.equal(outs[n][0], outs[1][0]) and torch.equal(outs[n][1], outs[1][1]), n


@pytest.mark.parametrize("C,T,H,W,ratio", [(128, 3, 40, 90, 2), (512, 2, 30, 45, 1), (256, 5, 24, 44, 4),
                                           (192, 3, 40, 77, 0), (128, 2, 64, 96, 0)])
def test_staged_norm_act_equals_the_register_fed_kernel(C, T, H, W, ratio):
    """tg_vae_norm_act on dense and on pitched outputs (same bits), against fp32 torch: partial runs at the row ends, 24 vectors
    per pixel (C = 192: idle lanes), odd T (first-frame rule of the zq up-sampling).  With the DEVELOPER library and
    TG_NORM_STAGED=1 the dense call takes the shared-memory-staged kernel (1-D bulk copies; not shipped, see csrc/vae.cu) and
    the pitched one the register-fed kernel — the same arithmetic, so still the same bits."""
    from tokensgen_b200 import _ext as E
    g = torch.Generator().manual_seed(C + W)
    x = (torch.randn(T, H, W, C, generator=g) * 1.5 + 0.3).bfloat16().cuda()
    gamma, beta = torch.randn(C, generator=g).bfloat16().cuda(), torch.randn(C, generator=g).bfloat16().cuda()
    groups = 32
    sums = E.vae_group_stats(x, groups)
    zy = zb = None
    spatial = ratio > 0
    if spatial:
        Tz = (T + 1) // 2 if T > 1 else 1
        table = torch.randn(Tz, H // ratio, W // ratio, 2 * C + 64, generator=g).bfloat16().cuda()
        zy, zb = table[..., :C], table[..., C + 64:]
    dense = torch.empty(T, H, W, C, device="cuda", dtype=torch.bfloat16)
    pitched = torch.empty(T, H, W, C + 8, device="cuda", dtype=torch.bfloat16)[..., :C]
    E.vae_norm_act(x, sums, groups, 1e-6, gamma, beta, dense, zy, zb, silu=True)
    E.vae_norm_act(x, sums, groups, 1e-6, gamma, beta, pitched, zy, zb, silu=True)
    assert torch.isfinite(dense.float()).all() and torch.equal(dense, pitched)
    # and against plain torch (fp32 GroupNorm over the whole tensor, nearest up-sampling of the tables)
    xf = x.float().permute(3, 0, 1, 2).unsqueeze(0)
    ref = torch.nn.functional.group_norm(xf, groups, gamma.float(), beta.float(), 1e-6)
    if spatial:
        def up(t_):
            t_ = t_.float().permute(3, 0, 1, 2).unsqueeze(0)
            if T > 1 and T % 2 == 1:
                first = torch.nn.functional.interpolate(t_[:, :, :1], size=(1, H, W))
                rest = torch.nn.functional.interpolate(t_[:, :, 1:], size=(T - 1, H, W))
                return torch.cat([first, rest], dim=2)
            return torch.nn.functional.interpolate(t_, size=(T, H, W))
        ref = ref * up(zy) + up(zb)
    ref = torch.nn.functional.silu(ref)[0].permute(1, 2, 3, 0)
    assert rel_l2(dense, ref) < 5e-3


def test_frames_to_rgb8_matches_host_postprocess_bit_exact():
    """K20: GPU uint8 pack == VideoProcessor.postprocess_video("np") followed by the exporter's
    (frame * 255).round().astype(uint8), including out-of-range and half-way values."""
    import numpy as np
    from tokensgen_b200 import _ext as E
    from tokensgen_b200.pipeline import VideoProcessor
    g = torch.Generator().manual_seed(3)
    video = (torch.randn(1, 3, 5, 24, 40, generator=g) * 0.8).bfloat16()
    video[0, 0, 0, 0, :8] = torch.tensor([-1.5, -1.0, 1.0, 1.5, 0.0, 1 / 255, -1 / 255, 0.00390625]).bfloat16()
    vp = VideoProcessor()
    ref = (vp.postprocess_video(video.cuda(), "np") * 255).round().astype(np.uint8)        # [B, F, H, W, 3]
    got = vp.postprocess_video(video.cuda(), "uint8")
    assert got.dtype == np.uint8 and got.shape == ref.shape
    assert np.array_equal(got, ref)
    planes = E.vae_frames_to_rgb8(video[0].cuda())
    assert np.array_equal(planes.cpu().numpy(), ref[0])


# ------------------------------------------------------------------------------------------------ full channel widths
FULL = dict(block_out_channels=(128, 256, 256, 512), latent_channels=16, layers_per_block=3, norm_num_groups=32,
            sample_height=480, sample_width=720, scaling_factor=0.7)


@pytest.fixture(scope="module")
def full_env():
    """The CogVideoX-5b VAE at its real widths (128, 256, 256, 512), 3 layers per block, 215.6 M parameters, seeded weights."""
    from oracle import vae as ov
    from oracle.synth import synth_state_dict
    from tokensgen_b200.vae import AutoencoderKLCogVideoX
    cfg = ov.VaeConfig(**FULL)
    sd = synth_state_dict(ov.vae_shapes(cfg), seed=2024)
    vae = AutoencoderKLCogVideoX(**FULL)
    vae.load_state_dict(sd, strict=True)
    vae = vae.to("cuda", torch.bfloat16).eval()
    return ov, cfg, {k: v.float() for k, v in sd.items()}, vae


def _zq_tables(zq):
    from tokensgen_b200 import vae as V
    Tz, hz, wz = zq.shape[2:]
    zq_cl = torch.zeros(Tz, hz, wz, 64, device="cuda", dtype=torch.bfloat16)
    zq_cl[..., :16] = cl(zq)
    return V._ZqTables(zq_cl)


def test_full_width_mid_block_resnet_on_a_60x90_latent_tile(full_env):
    """decoder.mid_block.resnets[0] at 512 channels (SpatialNorm -> SiLU -> 512->512 causal conv, twice, + identity) on a
    3-frame 60 x 90 latent tile, then a second 2-frame call that consumes the conv caches — the real mid-block geometry of a
    480 x 720 decode (autoencoder_kl_cogvideox.py:276-309, 148-188)."""
    ov, cfg, sd, vae = full_env
    from tokensgen_b200 import vae as V
    g = torch.Generator().manual_seed(8)
    name = "decoder.mid_block.resnets.0"
    blk = vae.decoder.mid_block.resnets[0]
    vae._clear_fake_context_parallel_cache()
    cache = {}
    for T in (3, 2):
        f = torch.randn(1, 512, T, 60, 90, generator=g).bfloat16()
        zq = torch.randn(1, 16, T, 60, 90, generator=g).bfloat16()
        ref = ov.resnet_block(sd, name, f.float(), zq.float(), cfg, cache)
        out = V._resnet(blk, cl(f), _zq_tables(zq)).x
        e = rel_l2(cf(out), ref)
        print(f"full-width mid-block resnet (512 ch, {T}x60x90): rel_l2 {e:.3e}")
        assert e < 5e-3, (T, e)
    vae._clear_fake_context_parallel_cache()


def test_full_width_last_up_block(full_env):
    """decoder.up_blocks[3] (the full-resolution block: 256 -> 128 with a 1x1 shortcut, then 3 x 128 -> 128) at its real
    widths on a 9-frame 64 x 96 tile with the latent at 1/8 resolution and 3 frames (zq up-sampling with the odd first
    frame), then a second 8-frame / 2-latent-frame call on the carried caches."""
    ov, cfg, sd, vae = full_env
    from tokensgen_b200 import vae as V
    g = torch.Generator().manual_seed(9)
    blk = vae.decoder.up_blocks[3]
    assert blk.upsamplers is None and [r.in_channels for r in blk.resnets] == [256, 128, 128, 128]
    vae._clear_fake_context_parallel_cache()
    cache = {}
    for T, Tz in ((9, 3), (8, 2)):
        f = torch.randn(1, 256, T, 64, 96, generator=g).bfloat16()
        zq = torch.randn(1, 16, Tz, 8, 12, generator=g).bfloat16()
        ref = f.float()
        for i in range(4):
            ref = ov.resnet_block(sd, f"decoder.up_blocks.3.resnets.{i}", ref, zq.float(), cfg, cache)
        zt = _zq_tables(zq)
        h = cl(f)
        for r in blk.resnets:
            h = V._resnet(r, h, zt, 32)      # the epilogue of every conv2 hands its output's group sums to the next norm1
        assert h.sums is not None and h.groups == 32
        h = h.x
        e = rel_l2(cf(h), ref)
        print(f"full-width up_blocks[3] (256->128->128->128->128, {T}x64x96): rel_l2 {e:.3e}")
        assert e < 1e-2, (T, e)          # 8 convolutions deep
    vae._clear_fake_context_parallel_cache()


def test_full_width_decode_and_encode_vs_oracle(full_env):
    """The whole full-width coder (4 up / down blocks x 4 / 3 resnets, 512-channel mid block) on a small frame: decode
    [1,16,3,8,12] -> [1,3,9,64,96] and encode [1,3,9,64,96] -> moments, against the fp32 oracle; PSNR of the decoded frames."""
    ov, cfg, sd, vae = full_env
    g = torch.Generator().manual_seed(10)
    vae.disable_tiling()
    z = torch.randn(1, 16, 3, 8, 12, generator=g).bfloat16()
    d = vae.decode(z.cuda()).sample
    refd = ov.decode(sd, cfg, z.float())
    x = (torch.rand(1, 3, 9, 64, 96, generator=g) * 2 - 1).bfloat16()
    m = vae.encode(x.cuda()).latent_dist.parameters
    refm = ov.encode(sd, cfg, x.float())
    ed, em, p = rel_l2(d, refd), rel_l2(m, refm), psnr(d, refd)
    print(f"full-width decode rel_l2 {ed:.3e} PSNR {p:.1f} dB; encode rel_l2 {em:.3e}")
    assert d.shape == refd.shape == (1, 3, 9, 64, 96) and m.shape == refm.shape
    assert ed < 1.5e-2 and em < 1.5e-2 and p >= 40.0


@pytest.mark.parametrize("cin,cout,groups,T,H,W,res", [
    (128, 128, 32, 2, 160, 250, True),     # tap-reuse pair kernel, 4 channels per group, ragged width, residual
    (128, 256, 32, 1, 150, 300, False),    # pair kernel with two channel tiles, 8 channels per group
    (256, 512, 32, 2, 9, 33, False),       # CTA-pair 256-wide tiles, 16 channels per group, odd number of pixel tiles
    (64, 64, 8, 3, 19, 70, True),          # single-CTA 64-wide tiles (tiny configs), 8 channels per group
    (64, 128, 2, 1, 37, 45, False),        # 64 channels per group (a chunk lies inside one group)
])
def test_conv_epilogue_group_statistics_equal_the_separate_pass(cin, cout, groups, T, H, W, res):
    """tg_conv_args.stats: the sums the epilogue accumulates are those of the STORED output — compared with tg_vae_group_stats
    over that output (fp64 sums of the same bf16 values: equal up to the fp32 partial-sum order) and with torch on the host."""
    from tokensgen_b200 import _ext as E
    from tokensgen_b200.vae import _PackedConv
    g = torch.Generator().manual_seed(cin + cout + W)
    conv = torch.nn.Conv3d(cin, cout, 3)
    with torch.no_grad():
        conv.weight.copy_(torch.randn(conv.weight.shape, generator=g) / (27 * cin) ** 0.5)
        conv.bias.copy_(torch.randn(cout, generator=g))
    conv = conv.to(torch.bfloat16)
    x = (torch.randn(1, cin, T + 2, H, W, generator=g) + 0.3).bfloat16()
    r = torch.randn(T, H, W, cout, generator=g).bfloat16().cuda() if res else None
    w, b = _PackedConv().get(conv.cuda())
    assert E.conv_stats_supported(cout, groups)
    stats = torch.zeros(2 * groups, device="cuda", dtype=torch.float64)
    xin = x[0].permute(1, 2, 3, 0).contiguous().cuda()
    y = E.vae_conv(xin, w, b, cout, 3, 3, 3, T, H, W, residual=r, stats=stats, stat_groups=groups)
    y_plain = E.vae_conv(xin, w, b, cout, 3, 3, 3, T, H, W, residual=r)
    torch.cuda.synchronize()
    assert torch.equal(y, y_plain)                                   # the statistics do not touch the output
    want = E.vae_group_stats(y, groups)
    yg = y.double().cpu().reshape(-1, groups, cout // groups)
    host = torch.cat([yg.sum(dim=(0, 2)), (yg * yg).sum(dim=(0, 2))])
    torch.cuda.synchronize()
    scale = host[groups:].abs().max().item()
    assert (stats.cpu() - host).abs().max().item() < 2e-5 * scale, (stats.cpu() - host).abs().max().item() / scale
    assert (want.cpu() - host).abs().max().item() < 1e-5 * scale
    # reproducible run to run: a CTA's partial sums do not depend on how its epilogue warps interleave (no fp32 atomics)
    again = torch.zeros_like(stats)
    E.vae_conv(xin, w, b, cout, 3, 3, 3, T, H, W, residual=r, stats=again, stat_groups=groups)
    torch.cuda.synchronize()
    assert torch.equal(again.float(), stats.float())
    # a second launch accumulates on top (the decoder's frame batches share nothing, but the contract is "+=")
    E.vae_conv(xin, w, b, cout, 3, 3, 3, T, H, W, residual=r, stats=stats, stat_groups=groups)
    torch.cuda.synchronize()
    assert (stats.cpu() - 2 * host).abs().max().item() < 4e-5 * scale


def test_fused_statistics_leave_the_coder_output_within_rounding(env):
    """Whole decoder / encoder with the statistics taken from the conv epilogues vs the separate statistics pass."""
    ov, cfg, sd, vae = env
    from tokensgen_b200 import vae as V
    g = torch.Generator().manual_seed(12)
    z = torch.randn(1, 16, 5, 12, 10, generator=g).bfloat16().cuda()
    x = (torch.rand(1, 3, 9, 96, 80, generator=g) * 2 - 1).bfloat16().cuda()
    vae.disable_tiling()
    outs = {}
    try:
        for fused in (True, False):
            V._FUSED_STATS = fused
            outs[fused] = (vae.decode(z).sample.clone(), vae.encode(x).latent_dist.parameters.clone())
    finally:
        V._FUSED_STATS = True
    # ~40 bf16 layers deep a 1e-7 difference in a statistic flips roundings and the two runs drift apart to the bf16 noise floor
    # (the distance each has from the fp32 oracle, 6-9e-3): equally good answers, not equal ones
    assert rel_l2(outs[True][0], outs[False][0]) < 1.2e-2 and rel_l2(outs[True][1], outs[False][1]) < 1.2e-2
    ref_d, ref_m = ov.decode(sd, cfg, z.float().cpu()), ov.encode(sd, cfg, x.float().cpu())
    for fused in (True, False):
        assert rel_l2(outs[fused][0], ref_d) < 1.5e-2 and rel_l2(outs[fused][1], ref_m) < 1.5e-2, fused
