"""Pins oracle/ against the golden vectors produced by the unmodified reference (oracle/make_goldens.py)."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

from oracle import dit as odit
from oracle import dpm as odpm
from oracle import fifo as ofifo
from oracle import rope as orope
from oracle.synth import dit_shapes, state_dict_digest, synth_state_dict

TINY = dict(heads=4, head_dim=64, layers=2, time_dim=128, text_dim=128, in_ch=16, out_ch=16, patch=2, vip_dim=128)


def sha(t):
    return hashlib.sha256(t.detach().cpu().contiguous().view(torch.uint8).numpy().tobytes()).hexdigest()


def rel_l2(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm()).item()


def tiny_cfg(use_vip):
    return odit.DitConfig(num_attention_heads=4, attention_head_dim=64, in_channels=16, out_channels=16, time_embed_dim=128,
                          text_embed_dim=128, num_layers=2, patch_size=2, vip_length=12, vip_embed_dim=128, vip_scale=0.6,
                          use_vip=use_vip)


def test_rope_tables_bit_exact(golden_dir):
    g = torch.load(os.path.join(golden_dir, "rope.pt"))
    cos, sin = orope.rope_3d(64, [[0, 0, 0], [3, 4, 6]], (3, 4, 6))
    assert torch.equal(cos, g["small_cos"]) and torch.equal(sin, g["small_sin"])
    cos, sin = orope.window_rope(64, 13, 30, 45)
    assert cos.shape == (17550, 64) and cos.dtype == torch.float32
    assert sha(cos) == g["full_cos_sha"] and sha(sin) == g["full_sin_sha"]
    assert torch.equal(cos[g["full_rows"]], g["full_cos_rows"])
    gt = np.array([1000, 1003.25, 1006.5, 1009.75, 1013], dtype=np.float32)
    cos, sin = orope.rope_3d_from_grids(64, gt, np.linspace(0, 30, 8, endpoint=False, dtype=np.float32),
                                        np.linspace(0, 45, 12, endpoint=False, dtype=np.float32))
    assert torch.equal(cos, g["cond_cos"]) and torch.equal(sin, g["cond_sin"])
    cos, sin = orope.rope_3d_from_grids(64, g["img_grid_t"].numpy(), np.arange(30, dtype=np.float32),
                                        np.arange(45, dtype=np.float32))
    assert sha(cos) == g["img_cos_sha"] and sha(sin) == g["img_sin_sha"]


def test_vip_grids_match_pipeline_formulas():
    (it, ih, iw), (ct, ch, cw) = orope.vip_grids(60, 90, 2, num_chunks=2, frames_per_chunk=13, vip_frames_per_chunk=4,
                                                 h_queries=8, w_queries=12, start_frame_idx=1000)
    assert it.tolist() == list(range(26)) and len(ih) == 30 and len(iw) == 45
    assert ct[:5].tolist() == [1000.0, 1003.25, 1006.5, 1009.75, 1013.0] and len(ct) == 12
    assert ch.tolist() == [0.0, 3.75, 7.5, 11.25, 15.0, 18.75, 22.5, 26.25]


@pytest.mark.parametrize("use_vip", [True, False])
def test_state_dict_layout_and_digest(golden_dir, use_vip):
    meta = json.load(open(os.path.join(golden_dir, "dit_tiny.json")))["vip" if use_vip else "plain"]
    shapes = dit_shapes(use_vip=use_vip, **TINY)
    assert shapes == meta["shapes"]  # key names + shapes of the instantiated reference model
    assert state_dict_digest(synth_state_dict(shapes, 1234)) == meta["digest"]


@pytest.mark.parametrize("use_vip", [True, False])
@pytest.mark.parametrize("per_frame", [True, False])
def test_dit_forward_matches_reference(golden_dir, use_vip, per_frame):
    g = torch.load(os.path.join(golden_dir, "dit_tiny.pt"))
    tag = ("vip" if use_vip else "plain") + ("_pf" if per_frame else "_ps")
    lat, text, vip, ts = g[tag + "_inputs"]
    sd = synth_state_dict(dit_shapes(use_vip=use_vip, **TINY), 1234)
    cfg = tiny_cfg(use_vip)
    y32 = odit.dit_forward(sd, cfg, lat, text, ts, vip, g["rope"], g["img_rope"], g["cond_rope"], torch.float32)
    assert y32.shape == g[tag + "_f32"].shape
    assert rel_l2(y32, g[tag + "_f32"]) < 2e-5
    y16 = odit.dit_forward(sd, cfg, lat, text, ts, vip, g["rope"], g["img_rope"], g["cond_rope"], torch.bfloat16)
    assert rel_l2(y16, g[tag + "_bf16"]) < 2e-2      # same op order as the reference, bf16 rounding noise only
    assert rel_l2(g[tag + "_bf16"], g[tag + "_f32"]) < 3e-2  # the band the reference's own bf16 run sits in


@pytest.mark.parametrize("use_vip", [True, False])
def test_block_matches_reference(golden_dir, use_vip):
    g = torch.load(os.path.join(golden_dir, "dit_tiny.pt"))
    hid, enc, temb, h_ref, e_ref = g[("vip" if use_vip else "plain") + "_block"]
    sd = synth_state_dict(dit_shapes(use_vip=use_vip, **TINY), 1234)
    h, e = odit.block_forward(sd, "transformer_blocks.1", tiny_cfg(use_vip), hid, enc, temb, g["rope"],
                              g["img_rope"], g["cond_rope"], torch.float32)
    assert rel_l2(h, h_ref) < 1e-5 and rel_l2(e, e_ref) < 1e-5


def test_dpm_tables_and_steps_bit_exact(golden_dir):
    g = torch.load(os.path.join(golden_dir, "dpm.pt"))
    tb = odpm.DpmTables()
    assert torch.equal(tb.alphas_cumprod, g["alphas_cumprod"]) and torch.equal(tb.betas, g["betas"])
    assert tb.trailing_timesteps(52).tolist() == g["timesteps52"].tolist()
    assert tb.trailing_timesteps(50).tolist() == g["timesteps50"].tolist()
    assert len(g["step_cases"]) >= 10
    for c in g["step_cases"]:
        # the goldens were produced by the reference ON THE CPU -> PyTorch's CPU scalar semantics (oracle/dpm.py:_smul)
        p, x0 = odpm.step(tb, c["model_output"], c["old"], c["t"], c["prev_t"], c["back"], c["sample"], c["n1"], c["n2"],
                          device_semantics="cpu")
        assert p.dtype == c["prev_sample"].dtype and torch.equal(p, c["prev_sample"])
        assert torch.equal(x0, c["x0"])
    x, n, y = g["renoise"]
    assert torch.equal(odpm.add_noise_to_xt(tb, x, n), y) and y.dtype == torch.float64


def test_fifo_schedule_matches_reference_controller(golden_dir):
    traces = json.load(open(os.path.join(golden_dir, "fifo_trace.json")))
    tb = odpm.DpmTables()
    t_tab, prev_tab, next_tab = ofifo.fifo_timestep_tables(tb.trailing_timesteps(52))
    for name, tr in traces.items():
        sched = ofifo.window_schedule(tr["num_frames"])
        flat = [w for wins in sched for w in wins]
        assert len(flat) == len(tr["records"])
        assert tr["emitted"] == tr["num_frames"]
        by_iter = {}
        for r in tr["records"]:
            by_iter.setdefault(r["iteration"], []).append(r)
        for it, wins in enumerate(sched):
            recs = sorted(by_iter.get(it, []), key=lambda r: r["start"])
            assert [(w.start, w.mid, w.end, w.real_end) for w in wins] == [(r["start"], r["mid"], r["end"], r["real_end"]) for r in recs]
            for w, r in zip(wins, recs):
                assert t_tab[w.start:w.end].tolist() == r["t"]
                assert prev_tab[w.start:w.end].tolist() == r["prev_t"]
                assert next_tab[w.start:w.end].tolist() == r["next_t"]
    # published counts: 1418 window forwards for edit.yaml (12 chunks), 195 iterations
    assert len(traces["edit"]["records"]) == 1418
    assert len(ofifo.window_schedule(12 * 13)) == 195
    assert sum(len(w) for w in ofifo.window_schedule(24 * 13)) == 2666
