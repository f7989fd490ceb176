"""Host logic of the base stage pinned to the reference's OWN pipeline run (tests/golden/pipeline_tiny.pt, produced by
oracle/make_goldens.py::gen_pipeline_tiny from the unmodified MPFIFOVideoIPAdapterCogVideoXPipeline.__call__ on tiny models):
everything that involves no model arithmetic — schedule, chunk arithmetic, position grids, the padding / CFG layout of the
condensed tokens, the noise-draw order and the diagonal FIFO capture — must agree exactly."""
import os

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def gold():
    return torch.load(os.path.join(ROOT, "tests", "golden", "pipeline_tiny.pt"), weights_only=False)


def _cpu_pipe(cfg):
    from tokensgen_b200.pipeline import MPFIFOVideoIPAdapterCogVideoXPipeline
    from tokensgen_b200.resampler import Resampler
    from tokensgen_b200.scheduler import CogVideoXDPMScheduler
    from tokensgen_b200.transformer import CogVideoXTransformer3DModel
    from tokensgen_b200.vae import AutoencoderKLCogVideoX
    with torch.device("meta"):
        dit = CogVideoXTransformer3DModel(**cfg["dit"])
        dit.set_vip_layers(None, length=18, func_type="1", scale=[0.6], resampler_params=cfg["resampler"])
        res = Resampler(**cfg["resampler"])
        vae = AutoencoderKLCogVideoX(**cfg["vae"])
    pipe = MPFIFOVideoIPAdapterCogVideoXPipeline(None, None, vae, dit, CogVideoXDPMScheduler.cogvideox_5b(), resampler=res)
    pipe._device = torch.device("cpu")
    return pipe, dit, res, vae


def test_mirror_state_dict_layouts_equal_the_reference_models(gold):
    _, dit, res, vae = _cpu_pipe(gold["config"])
    for name, m in (("dit", dit), ("resampler", res), ("vae", vae)):
        mine = {k: list(v.shape) for k, v in m.state_dict().items()}
        assert mine == gold["meta"][name]["shapes"], name


@pytest.mark.parametrize("flow", ["from_video", "from_tokens"])
def test_schedule_counts_and_grids(gold, flow):
    from tokensgen_b200.pipeline import retrieve_timesteps
    out, call = gold[flow], gold["config"]["call"]
    pipe = _cpu_pipe(gold["config"])[0]
    ts, n = retrieve_timesteps(pipe.scheduler, call["num_inference_steps"], "cpu", None)
    assert n == out["num_inference_steps"] and [int(t) for t in ts] == out["timesteps"].tolist()
    nf = (call["num_frames_per_chunk"] - 1) // pipe.vae_scale_factor_temporal + 1
    assert (nf, pipe.resampler.config.num_temporal_queries) == (out["nf_per_chunk"], out["vip_nf_per_chunk"])
    assert out["num_frames"] == 2 * nf                     # max_num_chunks_w_fifo=25 clamps to the 2 chunks given
    lat = torch.empty(1, nf, 16, call["height"] // 8, call["width"] // 8)
    img, cond, _, _ = pipe._vip_grids(lat, 2, nf, out["video_ipadapter_start_frame_idx"])
    for mine, ref in zip(img, out["vip_image_rotary_grid"]):
        assert np.array_equal(np.asarray(mine, np.float32), np.asarray(ref, np.float32))
    for mine, ref in zip(cond, out["vip_condition_rotary_grid"]):
        assert np.array_equal(np.asarray(mine, np.float32), np.asarray(ref, np.float32))


def test_priming_frame_is_the_first_draw_of_the_call_generator(gold):
    """pipeline_cogvideox_mp_fifo.py:1190: step 0 stores the PRE-step latent of the last frame, and it ends up LAST in
    fifo_latents; prepare_latents is the first consumer of `generator` (the VAE posterior uses the global one)."""
    from tokensgen_b200 import _ext as E
    init = E.randn_tensor((1, 3, 16, 60, 90), torch.Generator().manual_seed(gold["seeds"]["call"]), "cpu", torch.bfloat16)
    assert torch.equal(gold["from_tokens"]["fifo_latents"][:, -1], init[:, 2])
    assert torch.equal(gold["from_video"]["fifo_latents_last"], init[:, 2])
    old = gold["from_tokens"]["fifo_old_pred_original_sample"]
    assert len(old) == 12 and old[-1] is None and all(o is not None for o in old[:-1])
    assert tuple(gold["from_tokens"]["fifo_latents"].shape) == (1, 12, 16, 60, 90)


def test_condensed_token_padding_and_cfg_layout(gold):
    """:611-646 with image_embeddings given (gen.yaml): pad by repeating the last chunk's tokens, then [cond-for-uncond | cond]."""
    emb = gold["inputs"]["image_embeddings"]
    want = torch.cat([emb] + [emb[:, [-1]]] * (emb.shape[1] // 2), dim=1)
    want = torch.cat([want, want], dim=0)
    assert torch.equal(gold["from_tokens"]["image_embeddings"], want)
    ie = gold["from_video"]["image_embeddings"]                     # To2V flow: 2 chunks + the padded chunk, 2 queries each
    assert tuple(ie.shape) == (2, 6, 256, 2, 3) and torch.equal(ie[0], ie[1])
    pe = gold["from_tokens"]["prompt_embeds"]
    assert torch.equal(pe[0], gold["inputs"]["negative_prompt_embeds"][0].bfloat16())
    assert torch.equal(pe[1], gold["inputs"]["prompt_embeds"][0].bfloat16())
