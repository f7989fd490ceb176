"""GPU parity of the Resampler mirror (tokensgen_b200/resampler.py, through the C ABI) against the fp32 oracle restatement
of longvgen/video_ipadapter/resampler.py on the same bf16-rounded weights and inputs.  Tolerance: relative L2 <= 1e-2
(bf16 tensor-core chain of 2 Perceiver layers vs fp32; the reference's own bf16 run sits in the same band, see
tests/test_resampler_cpu.py)."""
import numpy as np
import pytest
import torch

from oracle import rope as orope
from oracle.resampler import ResamplerConfig, resampler_forward, resampler_shapes
from oracle.synth import synth_state_dict

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm()).item()


@pytest.mark.parametrize("geom", [dict(t=3, h=4, w=6, qt=2, qh=2, qw=3, B=2), dict(t=3, h=10, w=12, qt=2, qh=4, qw=6, B=1)])
def test_resampler_matches_oracle(geom):
    from tokensgen_b200.resampler import Resampler
    c = dict(dim=256, depth=2, dim_head=64, heads=4, num_height_queries=geom["qh"], num_width_queries=geom["qw"],
             num_temporal_queries=geom["qt"], embedding_dim=256, output_dim=256, max_height_seq_len=geom["h"],
             max_width_seq_len=geom["w"], max_temporal_seq_len=geom["t"])
    cfg = ResamplerConfig(**c)
    sd = synth_state_dict(resampler_shapes(cfg), seed=97)
    m = Resampler(**c)
    m.load_state_dict(sd, strict=True)
    m = m.to("cuda", torch.bfloat16).eval()
    lin = lambda a, b, n: np.linspace(a, b, n, endpoint=False, dtype=np.float32)
    image_rope = orope.rope_3d_from_grids(64, lin(0, geom["t"], geom["t"]), lin(0, geom["h"], geom["h"]), lin(0, geom["w"], geom["w"]))
    sampling_rope = orope.rope_3d_from_grids(64, lin(1000, 1000 + geom["t"], geom["qt"]), lin(0, geom["h"], geom["qh"]),
                                             lin(0, geom["w"], geom["qw"]))
    g = torch.Generator().manual_seed(5)
    x = torch.randn(geom["B"], geom["t"], geom["h"] * geom["w"], 256, generator=g).bfloat16()
    with torch.no_grad():
        y = m(x.cuda(), image_rotary_emb=image_rope, sampling_rotary_emb=sampling_rope)
    torch.cuda.synchronize()
    ref = resampler_forward(sd, cfg, x, image_rope, sampling_rope, torch.float32)
    assert y.shape == ref.shape
    err = rel_l2(y.float(), ref)
    print(f"resampler {geom}: rel_l2 {err:.3e}")
    assert err < 1e-2


def test_resampler_state_dict_keys_match_reference_layout():
    from tokensgen_b200.resampler import Resampler
    c = dict(dim=256, depth=2, dim_head=64, heads=4, num_height_queries=2, num_width_queries=3, num_temporal_queries=2,
             embedding_dim=256, output_dim=256)
    m = Resampler(**c)
    assert {k: list(v.shape) for k, v in m.state_dict().items()} == resampler_shapes(ResamplerConfig(**c))
