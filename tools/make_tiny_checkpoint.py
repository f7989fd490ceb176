"""Writes a tiny random-init checkpoint tree in the diffusers directory layout the CLI loads (infer_cogvideo_mp_fifo.py:
`<root>/CogVideoX/{transformer,vae,scheduler}`, `<root>/To2V/{vip.pt,resampler/}`), a short synthetic mp4 and a yaml config with
the reference's schema — everything `infer_cogvideo_mp_fifo.py --config <root>/tiny_edit.yaml` needs to run end to end on
one or more GPUs without the real weights.  usage: python tools/make_tiny_checkpoint.py <root>"""
import json
import os
import sys

import numpy as np
import torch
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def save_model(model, cfg, path, cls_name):
    from safetensors.torch import save_file
    os.makedirs(path, exist_ok=True)
    with open(os.path.join(path, "config.json"), "w") as f:
        json.dump(dict(cfg, _class_name=cls_name, _diffusers_version="0.31.0.dev0"), f)
    save_file({k: v.contiguous() for k, v in model.state_dict().items()}, os.path.join(path, "diffusion_pytorch_model.safetensors"))


def main(root):
    import cv2
    from tokensgen_b200.resampler import Resampler
    from tokensgen_b200.transformer import CogVideoXTransformer3DModel
    from tokensgen_b200.vae import AutoencoderKLCogVideoX
    torch.manual_seed(0)
    base, vip_dir = os.path.join(root, "CogVideoX"), os.path.join(root, "To2V")
    dit_cfg = dict(num_attention_heads=4, attention_head_dim=64, in_channels=16, out_channels=16, time_embed_dim=128,
                   text_embed_dim=128, num_layers=2, patch_size=2, use_rotary_positional_embeddings=True, attention_bias=True)
    dit = CogVideoXTransformer3DModel(**dit_cfg)
    rp = dict(dim=256, depth=1, dim_head=64, heads=4, num_height_queries=2, num_width_queries=3, num_temporal_queries=2,
              embedding_dim=256, output_dim=256, ff_mult=4, max_height_seq_len=6, max_width_seq_len=5, max_temporal_seq_len=3)
    vip_params = dict(image_encoder_path="", scale=[0.6], length=18, use_vae_as_encoder=True, func_type="1",
                      video_ipadapter_start_frame_idx=1000, resampler_params=rp)
    for p in dit.parameters():
        torch.nn.init.normal_(p, std=0.02)
    save_model(dit, dit_cfg, os.path.join(base, "transformer"), "CogVideoXTransformer3DModel")
    dit.set_vip_layers(None, **vip_params)
    for n, p in dit.named_parameters():
        if "vip_" in n:
            torch.nn.init.normal_(p, std=0.02)
    dit.save_vip_layers(vip_dir)                                            # -> To2V/vip.pt, like the reference's training script
    res = Resampler(**rp)
    save_model(res, rp, os.path.join(vip_dir, "resampler"), "Resampler")
    # 96 x 80: a size whose tile arithmetic is consistent (tile 48 x 40, overlaps 40 / 32 are multiples of 8), like 480 x 720
    vae_cfg = dict(block_out_channels=[64, 64, 64, 64], layers_per_block=1, norm_num_groups=8, sample_height=96, sample_width=80,
                   scaling_factor=0.7, latent_channels=16, temporal_compression_ratio=4)
    vae = AutoencoderKLCogVideoX(**vae_cfg)
    save_model(vae, vae_cfg, os.path.join(base, "vae"), "AutoencoderKLCogVideoX")
    os.makedirs(os.path.join(base, "scheduler"), exist_ok=True)
    with open(os.path.join(base, "scheduler", "scheduler_config.json"), "w") as f:
        json.dump(dict(_class_name="CogVideoXDDIMScheduler", beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
                       num_train_timesteps=1000, prediction_type="v_prediction", rescale_betas_zero_snr=True, snr_shift_scale=1.0,
                       timestep_spacing="trailing", set_alpha_to_one=True, clip_sample=False), f)
    # prompt embeddings stand in for the (third-party, out-of-scope) T5 encoder
    g = torch.Generator().manual_seed(1)
    torch.save({"prompt_embeds": torch.randn(1, 10, 128, generator=g), "negative_prompt_embeds": torch.randn(1, 10, 128, generator=g)},
               os.path.join(root, "prompt_embeds.pt"))
    # 40 frames of moving gradients, 80 x 96 (w x h), 10 fps
    wr = cv2.VideoWriter(os.path.join(root, "clip.mp4"), cv2.VideoWriter_fourcc(*"mp4v"), 10.0, (80, 96))
    yy, xx = np.mgrid[0:96, 0:80]
    for t in range(40):
        img = np.stack([(xx * 2 + 5 * t) % 256, (yy * 3 + 2 * t) % 256, (xx + yy + 7 * t) % 256], axis=-1).astype(np.uint8)
        wr.write(img)
    wr.release()
    cfg = dict(pretrained_model_name_or_path=base, name_prefix="tiny", output_dir=os.path.join(root, "outputs"), seed=42,
               num_inference_steps=12, num_frames_per_chunk=9, guidance_scale=6.0, use_separate_guidance=False, dtype="bf16",
               height=96, width=80, prompt_embeds_path=os.path.join(root, "prompt_embeds.pt"),
               use_2nd_stage=False, seed_2nd=42, longvgen_pca=None, sampling_mode="fifo",
               sampling_params=dict(lookahead_denoising=True, use_adaptive_padding=True, use_sliding_window_embedding=False, num_partitions=4),
               use_vip=True, pretrained_resampler_name_or_path=vip_dir, video_ipadapter_params=vip_params, use_lora=False, cache_idx=None,
               input_config=dict(public=dict(pad_to_fit=False, crop_to_fit=True, start_t=0, end_t=-1, sample_fps=10, output_fps=10,
                                             output_res=[96, 80], max_num_chunks_w_fifo=25, max_num_chunks_wo_fifo=1, num_videos_per_prompt=1),
                                 clip1=dict(prompt="moving gradients", video=os.path.join(root, "clip.mp4"), params=dict(max_num_chunks=4))))
    with open(os.path.join(root, "tiny_edit.yaml"), "w") as f:
        yaml.safe_dump(cfg, f, sort_keys=False)

    # ---- the T2To + To2V flow of config/infer/gen.yaml (use_2nd_stage): a patch-1 tokens transformer that generates the
    # 16 leading PCA coordinates of the condensed tokens, the normalisation statistics and the pickled PCA object
    from pca import PCA
    t2to_dir = os.path.join(root, "T2To")
    t2to_cfg = dict(dit_cfg, patch_size=1)
    t2to = CogVideoXTransformer3DModel(**t2to_cfg)
    for p in t2to.parameters():
        torch.nn.init.normal_(p, std=0.02)
    save_model(t2to, t2to_cfg, os.path.join(t2to_dir, "transformer"), "CogVideoXTransformer3DModel")
    g2 = torch.Generator().manual_seed(2)
    torch.save(PCA(None).fit(torch.randn(512, rp["output_dim"], generator=g2)), os.path.join(vip_dir, "pca.pt"))
    torch.save(torch.randn(1, 16, generator=g2) * 0.1, os.path.join(vip_dir, "mean.pt"))
    torch.save(torch.rand(1, 16, generator=g2) + 0.5, os.path.join(vip_dir, "std.pt"))
    gen = dict(cfg, name_prefix="tinygen", use_2nd_stage=True, pretrained_2nd_stage_model_name_or_path=t2to_dir,
               longvgen_mean=os.path.join(vip_dir, "mean.pt"), longvgen_std=os.path.join(vip_dir, "std.pt"),
               longvgen_pca=os.path.join(vip_dir, "pca.pt"), guidance_scale_2nd=6.0)
    gen["video_ipadapter_params"] = dict(vip_params, scale=[1.0])
    gen["input_config"] = dict(public=cfg["input_config"]["public"],
                               clip1=dict(prompt="moving gradients", params=dict(max_num_chunks=4)))
    with open(os.path.join(root, "tiny_gen.yaml"), "w") as f:
        yaml.safe_dump(gen, f, sort_keys=False)
    print("wrote", root)


if __name__ == "__main__":
    main(sys.argv[1])
