"""End-to-end plumbing of the pipeline mirrors on the GPU with tiny random-init models: conditioning video -> VAE encode ->
patch projection -> Resampler -> 12-step base clip with the diagonal FIFO capture -> FIFO stage -> clip-parallel decode,
and the T2To pipeline tail.  (Arithmetic parity of each stage has its own test file; here: shapes, bit-exact priming
capture, determinism, finiteness.)"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _tiny_pipe():
    from tokensgen_b200.pipeline import MPFIFOVideoIPAdapterCogVideoXPipeline
    from tokensgen_b200.resampler import Resampler
    from tokensgen_b200.scheduler import CogVideoXDPMScheduler
    from tokensgen_b200.transformer import CogVideoXTransformer3DModel
    from tokensgen_b200.vae import AutoencoderKLCogVideoX
    torch.manual_seed(0)
    dit = CogVideoXTransformer3DModel(num_attention_heads=4, attention_head_dim=64, time_embed_dim=128, text_embed_dim=128,
                                      num_layers=2, use_rotary_positional_embeddings=True, attention_bias=True)
    rp = dict(dim=256, depth=1, dim_head=64, heads=4, num_height_queries=2, num_width_queries=3, num_temporal_queries=2,
              embedding_dim=256, output_dim=256, max_height_seq_len=4, max_width_seq_len=6, max_temporal_seq_len=3)
    dit.set_vip_layers(None, length=18, func_type="1", scale=[0.6], resampler_params=rp)
    for p in dit.parameters():
        torch.nn.init.normal_(p, std=0.02)
    res = Resampler(**rp)
    vae = AutoencoderKLCogVideoX(block_out_channels=(64, 64, 64, 64), layers_per_block=1, norm_num_groups=8, sample_height=64,
                                 sample_width=96, scaling_factor=0.7)
    dev = torch.device("cuda")
    pipe = MPFIFOVideoIPAdapterCogVideoXPipeline(None, None, vae.to(dev, torch.bfloat16).eval(), dit.to(dev, torch.bfloat16).eval(),
                                                 CogVideoXDPMScheduler.cogvideox_5b(), resampler=res.to(dev, torch.bfloat16).eval())
    return pipe.to(dev)


def _run(pipe, seed):
    from tokensgen_b200.fifo import cogvideo_fifo_mp_v2
    g = torch.Generator().manual_seed(3)
    frames = torch.rand(1, 18, 3, 64, 96, generator=g) * 2 - 1
    pe = torch.randn(1, 10, 128, generator=g)
    ne = torch.randn(1, 10, 128, generator=g)
    # the conditioning clip's VAE posterior is sampled from the GLOBAL RNG in the reference (pipeline_cogvideox_mp_fifo.py:585), not
    # from the call's generator: pin it here so that reruns are comparable
    pipe.vae_posterior_generator = torch.Generator().manual_seed(1000 + seed)
    base = pipe(frames=frames, prompt_embeds=pe, negative_prompt_embeds=ne, height=64, width=96, num_frames_per_chunk=9,
                max_num_chunks=2, max_num_chunks_w_fifo=25, max_num_chunks_wo_fifo=1, num_inference_steps=12, guidance_scale=6.0,
                generator=torch.Generator().manual_seed(seed), vip_scale=[0.6], sampling_mode="fifo",
                sampling_params={"num_partitions": 4, "use_adaptive_padding": True}, cache_idx=None, output_type="np",
                return_dict=False)
    orig, video, cache = cogvideo_fifo_mp_v2([pipe], base, seed=seed)
    return base, orig, video


def test_base_stage_fifo_stage_and_decode():
    from tokensgen_b200 import _ext as E
    pipe = _tiny_pipe()
    base, orig, video = _run(pipe, 42)
    assert base.nf_per_chunk == 3 and base.vip_nf_per_chunk == 2 and base.num_frames == 6
    assert tuple(base.fifo_latents.shape) == (1, 12, 16, 8, 12) and len(base.fifo_old_pred_original_sample) == 12
    assert base.fifo_old_pred_original_sample[-1] is None and base.fifo_old_pred_original_sample[0] is not None
    # priming capture (pipeline_cogvideox_mp_fifo.py:1190): step 0 contributes the PRE-step latent of frame 2, stored last
    init = E.randn_tensor((1, 3, 16, 8, 12), torch.Generator().manual_seed(42), "cuda", torch.bfloat16)
    assert torch.equal(base.fifo_latents[:, -1], init[:, 2])
    assert tuple(base.image_embeddings.shape) == (2, 6, 256, 2, 3)   # CFG pair x (2 chunks + 1 pad) x 2 temporal queries
    assert video.shape == (1, 18, 64, 96, 3) and orig.shape == (1, 9, 64, 96, 3)
    assert np.isfinite(video).all() and np.isfinite(orig).all() and 0.0 <= video.min() and video.max() <= 1.0
    assert video.std() > 0
    # determinism: same seeds -> identical video
    _, _, video2 = _run(pipe, 42)
    assert np.array_equal(video, video2)
    _, _, video3 = _run(pipe, 43)
    assert not np.array_equal(video, video3)


def test_streaming_decode_under_the_fifo_loop_is_bit_identical():
    """`streaming_decode=True` decodes every chunk as soon as its frames have left the queue — in line on the
    denoising stream (the shipped default) or on a side stream beside the window forwards (TG_STREAM_DECODE_OVERLAP=1; fine at
    these shapes, off by default because of a full-size hang); the video equals the decode-after-the-loop path bit
    for bit, and each chunk was dispatched in its own iteration."""
    from tokensgen_b200 import fifo as F
    from tokensgen_b200.fifo import cogvideo_fifo_mp_v2
    pipe = _tiny_pipe()
    g = torch.Generator().manual_seed(3)
    frames = torch.rand(1, 27, 3, 64, 96, generator=g) * 2 - 1
    pe, ne = torch.randn(1, 10, 128, generator=g), torch.randn(1, 10, 128, generator=g)
    pipe.vae_posterior_generator = torch.Generator().manual_seed(5)
    base = pipe(frames=frames, prompt_embeds=pe, negative_prompt_embeds=ne, height=64, width=96, num_frames_per_chunk=9,
                max_num_chunks=3, max_num_chunks_w_fifo=25, max_num_chunks_wo_fifo=1, num_inference_steps=12, guidance_scale=6.0,
                generator=torch.Generator().manual_seed(42), vip_scale=[0.6], sampling_mode="fifo",
                sampling_params={"num_partitions": 4, "use_adaptive_padding": True}, cache_idx=None, output_type="np",
                return_dict=False)
    import copy
    _, ref, _ = cogvideo_fifo_mp_v2([pipe], copy.copy(base), seed=42)
    keep = F._DECODE_OVERLAP
    try:
        for overlap in (False, True):
            F._DECODE_OVERLAP = overlap
            log = {}
            _, got, _ = cogvideo_fifo_mp_v2([pipe], copy.copy(base), seed=42, streaming_decode=True, stream_log=log)
            assert got.shape == ref.shape == (1, 27, 64, 96, 3)
            assert np.array_equal(got, ref), overlap
            assert log == {c: (12 - 3) + 3 * (c + 1) - 1 for c in range(3)}      # T = 12, nf = 3
    finally:
        F._DECODE_OVERLAP = keep


def test_t2to_pipeline_tail():
    from pca import PCA
    from tokensgen_b200.pipeline_t2to import LongVGenCogVideoXPipeline
    from tokensgen_b200.scheduler import CogVideoXDPMScheduler
    from tokensgen_b200.transformer import CogVideoXTransformer3DModel
    torch.manual_seed(1)
    dit = CogVideoXTransformer3DModel(num_attention_heads=4, attention_head_dim=64, time_embed_dim=128, text_embed_dim=128,
                                      num_layers=2, patch_size=1, use_rotary_positional_embeddings=True, attention_bias=True)
    for p in dit.parameters():
        torch.nn.init.normal_(p, std=0.02)
    pipe = LongVGenCogVideoXPipeline(None, None, dit.to("cuda", torch.bfloat16).eval(), CogVideoXDPMScheduler.cogvideox_5b()).to("cuda")
    g = torch.Generator().manual_seed(2)
    pca = PCA(None).fit(torch.randn(64, 32, generator=g))
    mean, std = torch.randn(1, 32, generator=g), torch.rand(1, 32, generator=g) + 0.5
    out = pipe(prompt_embeds=torch.randn(1, 10, 128, generator=g), negative_prompt_embeds=torch.randn(1, 10, 128, generator=g),
               height=2, width=3, num_frames_per_chunk=2, num_chunks=4, num_inference_steps=4, guidance_scale=6.0,
               use_dynamic_cfg=True, generator=torch.Generator().manual_seed(5), longvgen_mean=mean, longvgen_std=std,
               longvgen_pca=pca).frames
    assert tuple(out.shape) == (1, 8, 32, 2, 3) and out.dtype == torch.bfloat16 and torch.isfinite(out.float()).all()


def test_fifo_stage_checkpoint_resume_is_bit_identical(tmp_path):
    """Restartable FIFO stage through the sampler entry point: crash after iteration 9, restart, resume from the state
    saved after iteration 8 (queue + x0 history + emitted frames; the condensed-token bookkeeping is fast-forwarded) —
    the latents equal the uninterrupted run's."""
    import copy
    from tokensgen_b200.fifo import cogvideo_fifo_mp_v2
    pipe = _tiny_pipe()
    g = torch.Generator().manual_seed(3)
    frames = torch.rand(1, 18, 3, 64, 96, generator=g) * 2 - 1
    pe, ne = torch.randn(1, 10, 128, generator=g), torch.randn(1, 10, 128, generator=g)
    base = pipe(frames=frames, prompt_embeds=pe, negative_prompt_embeds=ne, height=64, width=96, num_frames_per_chunk=9,
                max_num_chunks=2, max_num_chunks_w_fifo=25, max_num_chunks_wo_fifo=1, num_inference_steps=12, guidance_scale=6.0,
                generator=torch.Generator().manual_seed(42), vip_scale=[0.6], sampling_mode="fifo",
                sampling_params={"num_partitions": 4, "use_adaptive_padding": True}, cache_idx=None, output_type="latent",
                return_dict=False)
    _, ref, _ = cogvideo_fifo_mp_v2([pipe], copy.copy(base), seed=42)

    class Crash(Exception):
        pass

    def crash(it):
        if it == 9:
            raise Crash()

    ck = str(tmp_path / "ck")
    with pytest.raises(Crash):
        cogvideo_fifo_mp_v2([pipe], copy.copy(base), seed=42, checkpoint_dir=ck, checkpoint_every=4, progress=crash)
    import os
    names = sorted(os.listdir(ck))
    assert len(names) == 2 and names[0].endswith(".rank0.it000004.pt") and names[1].endswith(".rank0.it000008.pt")
    seen = []
    _, got, _ = cogvideo_fifo_mp_v2([pipe], copy.copy(base), seed=42, checkpoint_dir=ck, checkpoint_every=4, progress=seen.append)
    assert seen[0] == 8                      # resumed, not restarted
    assert torch.equal(got, ref)
    assert os.listdir(ck) == []              # the completed run removed its states
    # the same item with another seed must NOT resume from a state of the old run
    with pytest.raises(Crash):
        cogvideo_fifo_mp_v2([pipe], copy.copy(base), seed=42, checkpoint_dir=ck, checkpoint_every=4, progress=crash)
    seen2 = []
    cogvideo_fifo_mp_v2([pipe], copy.copy(base), seed=43, checkpoint_dir=ck, checkpoint_every=4, progress=seen2.append)
    assert seen2[0] == 0


def _golden_pipe():
    import os
    from oracle.synth import synth_state_dict
    from tokensgen_b200.pipeline import MPFIFOVideoIPAdapterCogVideoXPipeline
    from tokensgen_b200.resampler import Resampler
    from tokensgen_b200.scheduler import CogVideoXDPMScheduler
    from tokensgen_b200.transformer import CogVideoXTransformer3DModel
    from tokensgen_b200.vae import AutoencoderKLCogVideoX
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    gold = torch.load(os.path.join(root, "tests", "golden", "pipeline_tiny.pt"), weights_only=False)
    cfg, meta = gold["config"], gold["meta"]
    dit = CogVideoXTransformer3DModel(**cfg["dit"])
    dit.set_vip_layers(None, length=18, func_type="1", scale=[0.6], resampler_params=cfg["resampler"])
    res, vae = Resampler(**cfg["resampler"]), AutoencoderKLCogVideoX(**cfg["vae"])
    dev = torch.device("cuda")
    for name, m in (("dit", dit), ("resampler", res), ("vae", vae)):
        m.load_state_dict(synth_state_dict(meta[name]["shapes"], seed=meta[name]["seed"]), strict=True)
        m.to(dev, torch.bfloat16).eval()
    pipe = MPFIFOVideoIPAdapterCogVideoXPipeline(None, None, vae, dit, CogVideoXDPMScheduler.cogvideox_5b(), resampler=res).to(dev)
    return gold, pipe


_rel = lambda a, b: ((a.double().cpu() - b.double().cpu()).norm() / b.double().cpu().norm()).item()


def test_base_stage_against_the_reference_pipeline_golden():
    """The whole base stage (gen.yaml flow: condensed tokens given) against the reference's own
    MPFIFOVideoIPAdapterCogVideoXPipeline.__call__ run on CPU in bf16 (tests/golden/pipeline_tiny.pt): same deterministic
    weights, same prompt embeddings, same CPU generator -> identical noise draws; the 12-step CFG loop with the diagonal FIFO
    capture must agree within the bf16 tolerance, the arithmetic-free parts exactly."""
    gold, pipe = _golden_pipe()
    cfg = gold["config"]
    inp, ref = gold["inputs"], gold["from_tokens"]
    out = pipe(frames=None, image_embeddings=inp["image_embeddings"], prompt_embeds=inp["prompt_embeds"],
               negative_prompt_embeds=inp["negative_prompt_embeds"],
               generator=torch.Generator().manual_seed(gold["seeds"]["call"]), **cfg["call"])
    out = out[0] if isinstance(out, tuple) else out
    assert torch.equal(out.image_embeddings.cpu(), ref["image_embeddings"])
    assert torch.equal(out.fifo_latents[:, -1].cpu(), ref["fifo_latents"][:, -1])          # priming frame: pure noise draw
    e_fifo, e_orig = _rel(out.fifo_latents, ref["fifo_latents"]), _rel(out.orig_latents, ref["orig_latents"])
    print(f"base stage vs reference pipeline (bf16 CPU): fifo_latents rel_l2 {e_fifo:.3e}, orig_latents rel_l2 {e_orig:.3e}")
    assert e_fifo < 2e-2 and e_orig < 3e-2
    old = out.fifo_old_pred_original_sample
    assert len(old) == 12 and old[-1] is None
    assert _rel(old[0].reshape(ref["fifo_old_pred_original_sample"][0].shape), ref["fifo_old_pred_original_sample"][0]) < 3e-2


def test_from_video_flow_against_the_reference_pipeline_golden():
    """edit.yaml flow (pipeline_cogvideox_mp_fifo.py:562-648): conditioning video -> VAE encode (3 chunks of 9 frames at
    480 x 720, posterior SAMPLE x scaling) -> patch_embed.proj -> Resampler -> padding / CFG layout, against the condensed
    tokens of the reference's own run.  The reference draws the posterior noise from the global RNG (seeded in the generator
    script); the same CPU stream is replayed through `vae_posterior_generator`.  The call's `generator` must stay untouched by
    the encode: the priming frame (first draw of prepare_latents) and the grids are exact."""
    from oracle.synth import conditioning_clip
    gold, pipe = _golden_pipe()
    cfg, ref = gold["config"], gold["from_video"]
    pipe.vae_posterior_generator = torch.Generator().manual_seed(gold["seeds"]["global_rng"])
    inp = gold["inputs"]
    out = pipe(frames=conditioning_clip(18).to(torch.bfloat16), prompt_embeds=inp["prompt_embeds"],
               negative_prompt_embeds=inp["negative_prompt_embeds"],
               generator=torch.Generator().manual_seed(gold["seeds"]["call"]), **cfg["call"])
    out = out[0] if isinstance(out, tuple) else out
    ie = out.image_embeddings
    assert tuple(ie.shape) == tuple(ref["image_embeddings"].shape) and torch.equal(ie[0], ie[1])
    err = _rel(ie, ref["image_embeddings"])
    print(f"from-video flow: condensed tokens (VAE encode -> proj -> Resampler) vs the reference pipeline rel_l2 {err:.3e}")
    assert err < 3e-2
    assert torch.equal(out.fifo_latents[:, -1].cpu(), ref["fifo_latents_last"])
    assert (out.nf_per_chunk, out.vip_nf_per_chunk, out.num_frames) == (ref["nf_per_chunk"], ref["vip_nf_per_chunk"], ref["num_frames"])
    for mine, want in zip(out.vip_condition_rotary_grid, ref["vip_condition_rotary_grid"]):
        assert np.array_equal(np.asarray(mine, np.float32), np.asarray(want, np.float32))
