"""Oracle pin for the 3D causal VAE: oracle/vae.py vs golden vectors produced by the UNMODIFIED reference
AutoencoderKLCogVideoX (oracle/make_goldens.py:gen_vae_tiny).  fp32 on both sides -> tight tolerances."""
import os

import pytest
import torch

from oracle import vae as ov
from oracle.make_goldens import VAE_TINY
from oracle.synth import state_dict_digest, synth_state_dict


@pytest.fixture(scope="module")
def env(golden_dir):
    g = torch.load(os.path.join(golden_dir, "vae_tiny.pt"))
    cfg = ov.VaeConfig(**VAE_TINY)
    sd = synth_state_dict(ov.vae_shapes(cfg), seed=4321)
    assert state_dict_digest(sd) == g["digest"]
    return g, cfg, {k: v.float() for k, v in sd.items()}


def close(a, b, tol=2e-5):
    assert a.shape == b.shape, (a.shape, b.shape)
    err = ((a - b).norm() / b.norm()).item()
    assert err < tol, err


def test_frame_batches_match_reference_arithmetic():
    assert ov.frame_batches(49, 8, True) == [(0, 9), (9, 17), (17, 25), (25, 33), (33, 41), (41, 49)]
    assert ov.frame_batches(13, 2, False) == [(0, 3), (3, 5), (5, 7), (7, 9), (9, 11), (11, 13)]
    assert ov.frame_batches(1, 8, True) == [(0, 1)]
    assert ov.frame_batches(17, 8, True) == [(0, 9), (9, 17)]
    assert ov.frame_batches(5, 2, False) == [(0, 3), (3, 5)]


def test_untiled_encode_decode(env):
    g, cfg, sd = env
    close(ov.encode(sd, cfg, g["inputs"]["x"].float()), g["enc_f32"])
    close(ov.decode(sd, cfg, g["inputs"]["z"].float()), g["dec_f32"])
    # the reference's own bf16 run sits well inside the bf16 tolerance band used for the CUDA path
    band = ((g["dec_bf16"].float() - g["dec_f32"]).norm() / g["dec_f32"].norm()).item()
    assert band < 3e-2


def test_tiled_encode_decode(env):
    g, cfg, sd = env
    close(ov.decode(sd, cfg, g["inputs"]["zt"].float(), tiling=True), g["tiled_dec_f32"])
    close(ov.encode(sd, cfg, g["inputs"]["xt"].float(), tiling=True), g["tiled_enc_f32"])


def test_layers(env):
    g, cfg, sd = env
    a, b, ya, yb = g["conv_two_calls"]
    cache = {}
    name = "decoder.up_blocks.3.resnets.0.conv1"
    close(ov.causal_conv3d(sd, name, a, cache), ya)
    close(ov.causal_conv3d(sd, name, b, cache), yb)  # second call sees the first call's last two frames
    for T in (5, 2, 1):
        f, zq, y = g[f"spatial_norm_T{T}"]
        close(ov.spatial_norm(sd, "decoder.mid_block.resnets.0.norm1", f, zq, cfg.norm_num_groups, {}), y)
    for T in (3, 2, 1):
        h, y = g[f"upsample_time_T{T}"]
        close(ov.upsample3d(sd, "decoder.up_blocks.0.upsamplers.0", h, True), y)
    h, y = g["upsample_space"]
    close(ov.upsample3d(sd, "decoder.up_blocks.2.upsamplers.0", h, False), y)
    for T in (9, 8, 1):
        h, y = g[f"downsample_time_T{T}"]
        close(ov.downsample3d(sd, "encoder.down_blocks.0.downsamplers.0", h, True), y)
    h, y = g["downsample_space"]
    close(ov.downsample3d(sd, "encoder.down_blocks.2.downsamplers.0", h, False), y)


def test_posterior_sample_and_roundtrip_shapes(env):
    g, cfg, sd = env
    m = g["enc_f32"]
    eps = torch.randn(m.shape[0], m.shape[1] // 2, *m.shape[2:], generator=torch.Generator().manual_seed(1))
    z = ov.posterior_sample(m, eps)
    mean, logvar = m.chunk(2, dim=1)
    assert torch.equal(z, mean + torch.exp(0.5 * logvar.clamp(-30, 20)) * eps)
    assert z.shape == (1, 16, 5, 4, 5)          # 17 frames -> 5 latent frames, /8 spatial
    assert g["dec_f32"].shape == (1, 3, 17, 32, 40)  # 5 latent frames -> 17 frames


def test_tile_lanes_balance_the_tiled_coder_over_its_streams():
    """vae._tile_lanes: the 3 x 3 tiles of a 480 x 720 frame (latent 30x45 full tiles, 18-wide / 10-tall edge tiles) over 1 - 4
    stream lanes: every tile gets a lane, tile 0 stays on the caller's stream, and the lanes' areas differ by less than one full
    tile."""
    from tokensgen_b200.vae import _tile_lanes
    areas = [h * w for h in (30, 30, 10) for w in (45, 45, 18)]
    for n in (1, 2, 3, 4):
        lane = _tile_lanes(areas, n)
        assert len(lane) == 9 and lane[0] == 0 and set(lane) == set(range(n))
        load = [sum(a for a, l in zip(areas, lane) if l == k) for k in range(n)]
        assert sum(load) == sum(areas) and max(load) - min(load) < 30 * 45
    assert _tile_lanes([], 3) == [] and _tile_lanes([5], 3) == [0]
