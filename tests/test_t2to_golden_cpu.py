"""T2To stage pins (tests/golden/t2to_tiny.pt: the reference's LongVGenCogVideoXPipeline.__call__ and its patch_size = 1 DiT
run unmodified on the CPU — oracle/make_goldens.py::gen_t2to_tiny): RoPE tables with the 52/6/6 axis split
(pipeline_cogvideox_t2to.py:543-564) bit for bit, and the oracle's patch_size = 1 forward against the reference module."""
import os

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def gold():
    return torch.load(os.path.join(ROOT, "tests", "golden", "t2to_tiny.pt"), weights_only=False)


def test_rope_tables_52_6_6_are_bit_exact(gold):
    from oracle import rope as orope
    from tokensgen_b200.rope import get_3d_rotary_pos_embed_v2
    lin = lambda n: np.linspace(0, n, n, endpoint=False, dtype=np.float32)
    cos, sin = get_3d_rotary_pos_embed_v2(64, lin(8), lin(2), lin(3), dim_t=52, dim_h=6, dim_w=6)
    assert torch.equal(cos, gold["rope_cos"]) and torch.equal(sin, gold["rope_sin"])
    ocos, osin = orope.rope_3d_from_grids(64, lin(8), lin(2), lin(3), 52, 6, 6)
    assert torch.equal(ocos, gold["rope_cos"]) and torch.equal(osin, gold["rope_sin"])


def test_oracle_patch1_forward_equals_the_reference_module(gold):
    from oracle import dit as odit
    from oracle.synth import state_dict_digest, synth_state_dict
    sd = synth_state_dict(gold["meta"]["shapes"], seed=gold["seeds"]["dit"])
    assert state_dict_digest(sd) == gold["meta"]["digest"]
    cfg = odit.DitConfig(num_attention_heads=4, attention_head_dim=64, time_embed_dim=128, text_embed_dim=128, num_layers=2,
                         patch_size=1, use_vip=False)
    i = gold["inputs"]
    text = torch.cat([i["negative_prompt_embeds"], i["prompt_embeds"]])
    rope = (gold["rope_cos"], gold["rope_sin"])
    y = odit.dit_forward(sd, cfg, i["latents"], text, i["timestep"], None, rope, dtype=torch.float32)
    ref = gold["forward_f32"]
    assert ((y - ref).norm() / ref.norm()).item() < 1e-5
    # the reference's own bf16 run sits inside the band the GPU test allows
    assert ((gold["forward_bf16"].float() - ref).norm() / ref.norm()).item() < 1e-2


def test_mirror_ctor_signature_is_the_reference_one():
    """pipeline_cogvideox_t2to.py:297-311: (tokenizer, text_encoder, transformer, scheduler) — this stage has no VAE."""
    import inspect
    from tokensgen_b200.pipeline_t2to import LongVGenCogVideoXPipeline
    assert list(inspect.signature(LongVGenCogVideoXPipeline.__init__).parameters)[:5] == \
        ["self", "tokenizer", "text_encoder", "transformer", "scheduler"]
