"""Host-side contracts of the pipeline / CLI mirrors that need no GPU: config schema, PCA (the reference's only test:
pca.py vs sklearn), output-bundle fields, RoPE crop region, state-dict loading."""
import dataclasses
import json
import os

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_pca_matches_sklearn():
    """The reference's own self-check (pca.py:68-90): components, transform and inverse_transform vs sklearn on iris."""
    from sklearn import datasets
    from sklearn.decomposition import PCA as SkPCA

    from pca import PCA
    X = datasets.load_iris().data
    sk = SkPCA(n_components=2).fit(X)
    p = PCA(n_components=2).fit(torch.tensor(X))
    assert np.allclose(sk.components_, p.components_.numpy())
    assert np.allclose(sk.transform(X), p.transform(torch.tensor(X)).numpy())
    assert np.allclose(sk.inverse_transform(sk.transform(X)), p.inverse_transform(p.transform(torch.tensor(X))).numpy())
    # pickled whole, like weights/TokensGen-To2V/pca.pt
    import io
    buf = io.BytesIO()
    torch.save(p, buf)
    buf.seek(0)
    q = torch.load(buf, weights_only=False)
    assert torch.equal(q.components_, p.components_) and torch.equal(q.mean_, p.mean_)


@pytest.mark.parametrize("name", ["edit", "gen"])
def test_shipped_configs_follow_the_reference_schema(name):
    from tokensgen_b200 import config as c
    cfg = c.load(os.path.join(ROOT, "config", "infer", f"{name}.yaml"))
    assert cfg.num_inference_steps == 52 and cfg.num_frames_per_chunk == 49 and cfg.dtype == "bf16"
    assert cfg.sampling_params.num_partitions * 13 == cfg.num_inference_steps       # structural requirement of the FIFO
    rp = cfg.video_ipadapter_params.resampler_params
    assert cfg.video_ipadapter_params.length == (rp.num_temporal_queries + 1) * rp.num_height_queries * rp.num_width_queries == 480
    assert cfg.get("sampling_mode") == "fifo" and cfg.get("missing", 7) == 7
    pub = cfg.input_config.pop("public")
    for item in cfg.input_config.values():
        dps = __import__("copy").deepcopy(pub)
        dps.update(item.get("params", {}))
        assert dps.max_num_chunks in (12, 24) and dps.output_res == [480, 720] and dps.output_fps == 10
    assert cfg.use_2nd_stage == (name == "gen")


def test_output_bundle_fields_match_reference():
    """FIFOCogVideoXPipelineOutput field names/order (pipeline_cogvideox_mp_fifo.py:268-296) — the sampler reads them by name."""
    from tokensgen_b200.pipeline import FIFOCogVideoXPipelineOutput
    ref = ["fifo_latents", "fifo_old_pred_original_sample", "orig_latents", "nf_per_chunk", "vip_nf_per_chunk", "num_frames",
           "image_embeddings", "timesteps", "num_inference_steps", "do_classifier_free_guidance", "use_separate_guidance",
           "use_dynamic_cfg", "prompt_embeds", "image_rotary_emb", "vip_image_rotary_grid", "vip_condition_rotary_grid",
           "cache_idx", "attention_kwargs", "guidance_scale", "guidance_scale_img", "extra_step_kwargs", "condition_frames",
           "video_ipadapter_start_frame_idx", "sampling_params", "output_type", "return_dict"]
    assert [f.name for f in dataclasses.fields(FIFOCogVideoXPipelineOutput)] == ref


def test_pipeline_call_signature_keeps_reference_keywords():
    import inspect

    from tokensgen_b200.pipeline import MPFIFOVideoIPAdapterCogVideoXPipeline as P
    names = list(inspect.signature(P.__call__).parameters)
    for k in ("prompt", "frames", "image_embeddings", "num_inference_steps", "num_frames_per_chunk", "max_num_chunks",
              "max_num_chunks_w_fifo", "max_num_chunks_wo_fifo", "guidance_scale", "guidance_scale_img", "use_separate_guidance",
              "generator", "vip_scale", "sampling_mode", "sampling_params", "cache_idx", "video_ipadapter_start_frame_idx",
              "return_dict", "prompt_embeds", "negative_prompt_embeds", "latents"):
        assert k in names
    assert hasattr(P, "preprare_for_fifo") and hasattr(P, "decode_latents") and hasattr(P, "_prepare_vip_rotary_positional_embeddings")


def test_rope_crop_region_full_frame():
    from tokensgen_b200.pipeline import get_resize_crop_region_for_grid
    assert get_resize_crop_region_for_grid((30, 45), 45, 30) == ((0, 0), (30, 45))
    (t, l), (b, r) = get_resize_crop_region_for_grid((4, 6), 45, 30)
    assert (b - t, r - l) == (30, 45)


def test_from_pretrained_roundtrip(tmp_path):
    """diffusers directory layout: config.json + safetensors -> mirror with identical tensors (Resampler as the small case)."""
    from safetensors.torch import save_file

    from tokensgen_b200.resampler import Resampler
    cfg = dict(dim=256, depth=1, dim_head=64, heads=4, num_height_queries=2, num_width_queries=3, num_temporal_queries=2,
               embedding_dim=256, output_dim=256, max_height_seq_len=4, max_width_seq_len=6, max_temporal_seq_len=3)
    m = Resampler(**cfg)
    d = tmp_path / "resampler"
    d.mkdir()
    (d / "config.json").write_text(json.dumps(dict(cfg, _class_name="Resampler", _diffusers_version="0.31.0.dev0")))
    save_file({k: v.contiguous() for k, v in m.state_dict().items()}, str(d / "diffusion_pytorch_model.safetensors"))
    m2 = Resampler.from_pretrained(str(tmp_path), subfolder="resampler", torch_dtype=torch.bfloat16)
    assert m2.dtype == torch.bfloat16 and m2.config.heads == 4
    for k, v in m.state_dict().items():
        assert torch.equal(m2.state_dict()[k], v.to(torch.bfloat16))
    # meta construction + assignment: every tensor materialised on the requested device, nothing left on `meta`
    m3 = Resampler.from_pretrained(str(tmp_path), subfolder="resampler", torch_dtype=torch.bfloat16, device="cpu")
    assert all(not t.is_meta and t.device.type == "cpu" for t in list(m3.parameters()) + list(m3.buffers()))
    # a checkpoint of another geometry is refused (assignment would otherwise accept any shape) ...
    sd = {k: v.contiguous() for k, v in m.state_dict().items()}
    k0 = next(k for k, v in sd.items() if v.dim() == 2)
    save_file(dict(sd, **{k0: sd[k0][:-1].contiguous()}), str(d / "diffusion_pytorch_model.safetensors"))
    with pytest.raises(RuntimeError, match="size mismatch"):
        Resampler.from_pretrained(str(tmp_path), subfolder="resampler", torch_dtype=torch.bfloat16)
    # ... and so is one that lacks tensors
    save_file({k: v for k, v in sd.items() if k != k0}, str(d / "diffusion_pytorch_model.safetensors"))
    with pytest.raises(RuntimeError, match="lacks"):
        Resampler.from_pretrained(str(tmp_path), subfolder="resampler", torch_dtype=torch.bfloat16)


def test_scheduler_from_config_and_reference_defaults():
    """a partial config must mean what it means in the reference (scheduling_dpm_cogvideox.py:181-197 defaults:
    epsilon / leading / no zero-SNR rescale / snr_shift_scale 3.0), diffusers bookkeeping keys are dropped, anything else
    unknown is an error."""
    import inspect
    from tokensgen_b200.scheduler import COGVIDEOX_5B_CONFIG, CogVideoXDPMScheduler
    sig = inspect.signature(CogVideoXDPMScheduler.__init__).parameters
    assert (sig["prediction_type"].default, sig["timestep_spacing"].default, sig["rescale_betas_zero_snr"].default,
            sig["snr_shift_scale"].default, sig["clip_sample"].default) == ("epsilon", "leading", False, 3.0, True)
    s = CogVideoXDPMScheduler.from_config({"_class_name": "CogVideoXDDIMScheduler", "_diffusers_version": "0.31.0.dev0",
                                           **COGVIDEOX_5B_CONFIG, "timestep_spacing": "leading"}, timestep_spacing="trailing")
    assert s.config.timestep_spacing == "trailing"
    s.set_timesteps(52)
    assert int(s.timesteps[0]) == 999 and len(s.timesteps) == 52
    with pytest.raises(NotImplementedError, match="v_prediction"):
        CogVideoXDPMScheduler.from_config({"beta_end": 0.012})          # prediction_type missing -> the reference's "epsilon"
    with pytest.raises(ValueError, match="unknown keys"):
        CogVideoXDPMScheduler.from_config({**COGVIDEOX_5B_CONFIG, "thresholding": True})
    a, b = CogVideoXDPMScheduler.cogvideox_5b(), CogVideoXDPMScheduler.cogvideox_5b(snr_shift_scale=3.0)
    assert not torch.equal(a.alphas_cumprod, b.alphas_cumprod)
