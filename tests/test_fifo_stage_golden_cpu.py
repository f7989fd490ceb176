"""FIFO stage pinned on TENSORS to the reference sampler itself (tests/golden/fifo_stage_tiny.pt: the unmodified
`cogvideo_fifo_mp_v2` driving its own worker `fifo_onestep_per_gpu` with the reference DiT / scheduler on CPU, noise replaced
by oracle.synth.keyed_noise — oracle/make_goldens.py::gen_fifo_stage).

CPU half: the controller.  Our `run_fifo` is driven with a step function that REPLAYS the reference worker's recorded
outputs; everything the controller does with tensors — queue priming, window slicing, x0-history bookkeeping, lookahead
write-back, emit, shift + re-noise, the condensed-token lookup (embedding slice, position grids) — must then reproduce the
reference's window inputs and its final latents bit for bit.  (The GPU half, tests/test_fifo_stage_gpu.py, runs the real
window step.)"""
import os

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def gold():
    return torch.load(os.path.join(ROOT, "tests", "golden", "fifo_stage_tiny.pt"), weights_only=False)


def build_stage(gold, device="cpu"):
    """Schedule / queue / vip book of the golden's priming state, through the product's own classes."""
    from oracle.make_goldens import fifo_tiny_base_output
    from tokensgen_b200.fifo import FifoQueue, FifoSchedule, VipBook
    b = fifo_tiny_base_output(device)
    G = gold["config"]["geom"]
    sched = FifoSchedule(b["num_frames"], [int(t) for t in gold["timesteps"]], b["nf_per_chunk"], G["num_partitions"], True)
    queue = FifoQueue(b["fifo_latents"], b["fifo_old_pred_original_sample"], sched.r_nf)
    vip = VipBook(b["vip_image_rotary_grid"], b["vip_condition_rotary_grid"], b["image_embeddings"], b["nf_per_chunk"],
                  b["vip_nf_per_chunk"], G["T"], b["video_ipadapter_start_frame_idx"])
    return b, sched, queue, vip


def test_controller_replay_reproduces_the_reference_sampler_bit_for_bit(gold):
    from oracle import dpm as odpm
    from oracle.synth import keyed_noise
    from tokensgen_b200.fifo import run_fifo
    b, sched, queue, vip = build_stage(gold)
    calls = {(c["it"], c["start"]): c for c in gold["calls"]}
    tables = odpm.DpmTables()
    seen, checked = [], []

    def step_fn(w, lat, old, t, pt, nt, gen):
        rec = calls[(w.iteration, w.start)]
        seen.append((w.iteration, w.start))
        assert (w.mid, w.end, w.real_end) == (rec["mid"], rec["end"], rec["real_end"])
        if "lat_in" in rec:                       # iterations 0 / 7 / 14: the window inputs themselves
            assert torch.equal(lat, rec["lat_in"])
            assert [o is None for o in old] == [o is None for o in rec["old_in"]]
            assert all(o is None or torch.equal(o.reshape(r.shape), r) for o, r in zip(old, rec["old_in"]))
            img, cond, emb, _ = vip.window(w.start, w.end)
            assert np.array_equal(img[0], rec["img_t"]) and np.array_equal(cond[0], rec["cond_t"])
            assert torch.equal(emb, rec["emb_in"])
            checked.append(w.iteration)
        return rec["lat_out"], [rec["x0_out"][:, [j]] for j in range(rec["x0_out"].shape[1])]

    def shift_fn(q, gen):                        # shift_latents (cogvideo_sampling_mp_fifo.py:117-131) on the CPU
        it = seen[-1][0]
        q.latents[:, :-1] = q.latents[:, 1:].clone()
        q.x0[:-1] = q.x0[1:].clone()
        q.x0_valid = q.x0_valid[1:] + [False]
        q.latents[:, -1] = odpm.add_noise_to_xt(tables, q.latents[:, -1], keyed_noise((it, 999, 0, 0), q.latents[:, -1].shape))
        vip.shift()

    emitted = run_fifo(sched, queue, step_fn, shift_fn, seed=0)
    assert sorted(seen) == sorted(calls) and len(seen) == len(gold["calls"])     # same window calls, none extra
    assert sorted(set(checked)) == [0, 7, 14]
    video = torch.cat(emitted[gold["config"]["geom"]["T"] - b["nf_per_chunk"]:], dim=1)
    assert torch.equal(video, gold["video"])
    assert torch.equal(b["orig_latents"], gold["orig"])


def test_reference_worker_agrees_with_the_oracle_window_step(gold):
    """The recorded reference worker outputs against the oracle restatement (DiT forward op by op in bf16 + the CPU-semantics
    scheduler chain) on the recorded inputs of iteration 7: pins oracle.dit / oracle.dpm to the reference WORKER
    (cogvideo_sampling_mp_fifo.py:408-579), vip-RoPE rebuild and per-frame timesteps included."""
    from oracle import dit as odit
    from oracle import dpm as odpm
    from oracle import rope as orope
    from oracle.make_goldens import fifo_tiny_base_output
    from oracle.synth import keyed_noise, synth_state_dict
    b = fifo_tiny_base_output()
    c = gold["config"]
    sd = synth_state_dict(gold["meta"]["shapes"], seed=gold["seeds"]["dit"])
    cfg = odit.DitConfig(num_attention_heads=4, attention_head_dim=64, time_embed_dim=128, text_embed_dim=128, num_layers=2,
                         vip_length=c["vip"]["length"], vip_embed_dim=c["geom"]["vip_dim"], use_vip=True)
    nf, gh, gw = b["rope_grid"]
    rope = orope.rope_3d(64, [[0, 0, 0], [nf, gh, gw]], (nf, gh, gw))
    from tokensgen_b200.fifo import FifoSchedule
    sched = FifoSchedule(b["num_frames"], [int(t) for t in gold["timesteps"]], nf, c["geom"]["num_partitions"], True)
    tables = odpm.DpmTables()
    worst = 0.0
    for rec in [r for r in gold["calls"] if r["it"] == 7][:3]:
        s, e = rec["start"], rec["end"]
        img = orope.rope_3d_from_grids(64, rec["img_t"], b["vip_image_rotary_grid"][1], b["vip_image_rotary_grid"][2])
        cond = orope.rope_3d_from_grids(64, rec["cond_t"], b["vip_condition_rotary_grid"][1], b["vip_condition_rotary_grid"][2])
        ts = torch.as_tensor(sched.t[s:e].copy()).expand(2, -1)
        lat = rec["lat_in"]
        npred = odit.dit_forward(sd, cfg, torch.cat([lat, lat]), b["prompt_embeds"], ts, rec["emb_in"], rope, img, cond, torch.bfloat16)
        n1 = torch.cat([keyed_noise((7, s, j, 0), lat[:, [j]].shape) for j in range(nf)], dim=1)
        n2 = torch.cat([keyed_noise((7, s, j, 1), lat[:, [j]].shape) for j in range(nf)], dim=1)
        out, x0s = odpm.window_step_bf16(tables, npred.bfloat16(), c["guidance_scale"], lat, rec["old_in"], sched.t[s:e],
                                         sched.prev_t[s:e], sched.next_t[s:e], n1, n2, device_semantics="cpu")
        err = ((out.float() - rec["lat_out"].float()).norm() / rec["lat_out"].float().norm()).item()
        ex0 = ((torch.cat(x0s, 1).float() - rec["x0_out"].float()).norm() / rec["x0_out"].float().norm()).item()
        worst = max(worst, err, ex0)
    assert worst < 1e-2, worst


def test_separate_guidance_worker_agrees_with_the_oracle(gold):
    """use_separate_guidance (three branches uncond_txt / uncond_img / txt_img, cogvideo_sampling_mp_fifo.py:493-497,528-530):
    the reference worker's recorded outputs on the windows of iteration 7 against the oracle's B = 3 forward + 3-branch CFG."""
    from oracle import dit as odit
    from oracle import dpm as odpm
    from oracle import rope as orope
    from oracle.make_goldens import fifo_tiny_base_output
    from oracle.synth import keyed_noise, synth_state_dict
    from tokensgen_b200.fifo import FifoSchedule
    b = fifo_tiny_base_output()
    c = gold["config"]
    sd = synth_state_dict(gold["meta"]["shapes"], seed=gold["seeds"]["dit"])
    cfg = odit.DitConfig(num_attention_heads=4, attention_head_dim=64, time_embed_dim=128, text_embed_dim=128, num_layers=2,
                         vip_length=c["vip"]["length"], vip_embed_dim=c["geom"]["vip_dim"], use_vip=True)
    nf, gh, gw = b["rope_grid"]
    rope = orope.rope_3d(64, [[0, 0, 0], [nf, gh, gw]], (nf, gh, gw))
    sched = FifoSchedule(b["num_frames"], [int(t) for t in gold["timesteps"]], nf, c["geom"]["num_partitions"], True)
    tables = odpm.DpmTables()
    recs = [r for r in gold["calls"] if "sep_lat_out" in r]
    assert len(recs) >= 4
    for rec in recs[:2]:
        s, e = rec["start"], rec["end"]
        img = orope.rope_3d_from_grids(64, rec["img_t"], b["vip_image_rotary_grid"][1], b["vip_image_rotary_grid"][2])
        cond = orope.rope_3d_from_grids(64, rec["cond_t"], b["vip_condition_rotary_grid"][1], b["vip_condition_rotary_grid"][2])
        ts = torch.as_tensor(sched.t[s:e].copy()).expand(3, -1)
        lat = rec["lat_in"]
        npred = odit.dit_forward(sd, cfg, torch.cat([lat] * 3), b["prompt_embeds_sep"], ts, rec["sep_emb_in"], rope, img, cond,
                                 torch.bfloat16)
        n1 = torch.cat([keyed_noise((7, s, j, 0), lat[:, [j]].shape) for j in range(nf)], dim=1)
        n2 = torch.cat([keyed_noise((7, s, j, 1), lat[:, [j]].shape) for j in range(nf)], dim=1)
        out, x0s = odpm.window_step_bf16(tables, npred.bfloat16(), c["guidance_scale"], lat, rec["old_in"], sched.t[s:e],
                                         sched.prev_t[s:e], sched.next_t[s:e], n1, n2, device_semantics="cpu",
                                         guidance_scale_img=b["guidance_scale_img"])
        rel = lambda a, r: ((a.float() - r.float()).norm() / r.float().norm()).item()
        assert rel(out, rec["sep_lat_out"]) < 1e-2 and rel(torch.cat(x0s, 1), rec["sep_x0_out"]) < 1e-2
        assert rel(rec["sep_x0_out"], rec["x0_out"]) > 5e-2            # the three-branch result is a different one
