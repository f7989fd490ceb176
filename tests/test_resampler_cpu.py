"""Pins oracle/resampler.py against the golden produced by the unmodified reference Resampler
(longvgen/video_ipadapter/resampler.py, run through oracle/stubs by oracle/make_goldens.py::gen_resampler_tiny)."""
import os

import torch

from oracle.make_goldens import RESAMPLER_TINY
from oracle.resampler import ResamplerConfig, resampler_forward, resampler_shapes
from oracle.synth import state_dict_digest, synth_state_dict


def rel_l2(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm()).item()


def test_resampler_oracle_matches_reference(golden_dir):
    g = torch.load(os.path.join(golden_dir, "resampler_tiny.pt"))
    cfg = ResamplerConfig(**RESAMPLER_TINY)
    sd = synth_state_dict(resampler_shapes(cfg), seed=2468)
    assert state_dict_digest(sd) == g["digest"]
    y = resampler_forward(sd, cfg, g["x"], g["image_rope"], g["sampling_rope"], torch.float32)
    assert y.shape == g["out_f32"].shape == (2, 2, 128, 2, 3)
    assert rel_l2(y, g["out_f32"]) < 1e-5  # same fp32 ops in the same order; only SDPA backend choice may differ
    yb = resampler_forward(sd, cfg, g["x"], g["image_rope"], g["sampling_rope"], torch.bfloat16)
    assert rel_l2(yb, g["out_bf16"]) < 2e-2   # bf16 chains: the band the reference's own bf16 run sits in vs fp32
    assert rel_l2(g["out_bf16"], g["out_f32"]) < 2e-2


def test_resampler_shapes_full_config():
    cfg = ResamplerConfig()  # config/infer/edit.yaml:45-58
    s = resampler_shapes(cfg)
    assert s["latents"] == [1, 384, 3072] and s["layers.3.0.to_kv.weight"] == [2048, 3072]
    assert sum(int(torch.tensor(v).prod()) for v in s.values()) > 300e6  # ~0.35 B parameters
