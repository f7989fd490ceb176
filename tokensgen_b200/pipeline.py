"""Host-side mirror of the reference's To2V pipeline, `MPFIFOVideoIPAdapterCogVideoXPipeline`
(longvgen/pipeline/pipeline_cogvideox_mp_fifo.py:268-1518): prompt embeddings, VAE-encode of the conditioning video ->
patch projection -> Resampler (condensed tokens), the 52-step base denoise of the first clip that also captures the
diagonal FIFO priming state, `decode_latents`, the RoPE / vip-RoPE grid builders, and `preprare_for_fifo` (sic).

Same class name, `from_pretrained(path, transformer=, resampler=, torch_dtype=)`, `__call__` keyword arguments and
`FIFOCogVideoXPipelineOutput` fields, so `infer_cogvideo_mp_fifo.py` drives it unchanged.  All model arithmetic runs in the
CUDA mirrors (transformer / resampler / vae / scheduler); this file is orchestration.

Deliberate differences (results unchanged):
  * with classifier-free guidance and `use_separate_guidance=False` the reference VAE-encodes (num_chunks+1) all-zero clips
    and runs the Resampler on each, then uses the result ONLY for its length (:618-646).  Here the length is computed
    (`(num_chunks + 1) * num_temporal_queries`) and the 2 x 25 x 149 TFLOP of dead encodes are skipped.
  * the text encoder (T5) is a third-party model outside the hot path: pass `prompt_embeds` / `negative_prompt_embeds`,
    or give the pipeline a `transformers` T5 encoder + tokenizer and it is called exactly like `_get_t5_prompt_embeds`.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass, field
from types import SimpleNamespace
from typing import Any, Dict, List, Optional, Tuple, Union

import numpy as np
import torch

from . import _ext as E
from .rope import get_3d_rotary_pos_embed, get_3d_rotary_pos_embed_v2
from .scheduler import CogVideoXDPMScheduler


@dataclass
class FIFOCogVideoXPipelineOutput:
    """pipeline_cogvideox_mp_fifo.py:268-296 (same field names and order)."""
    fifo_latents: torch.Tensor
    fifo_old_pred_original_sample: List[Optional[torch.Tensor]]
    orig_latents: torch.Tensor
    nf_per_chunk: int
    vip_nf_per_chunk: Optional[int]
    num_frames: int
    image_embeddings: Optional[torch.Tensor]
    timesteps: torch.Tensor
    num_inference_steps: int
    do_classifier_free_guidance: bool
    use_separate_guidance: bool
    use_dynamic_cfg: bool
    prompt_embeds: torch.Tensor
    image_rotary_emb: Optional[Tuple[torch.Tensor, torch.Tensor]]
    vip_image_rotary_grid: Optional[List[np.ndarray]]
    vip_condition_rotary_grid: Optional[List[np.ndarray]]
    cache_idx: List[int]
    attention_kwargs: Optional[Dict[str, Any]] = None
    guidance_scale: float = 6
    guidance_scale_img: float = 6
    extra_step_kwargs: Optional[Dict[str, Any]] = None
    condition_frames: Optional[torch.Tensor] = None
    video_ipadapter_start_frame_idx: Optional[int] = 1000
    sampling_params: Dict[str, Any] = None
    output_type: str = "pil"
    return_dict: bool = True


@dataclass
class CogVideoXPipelineOutput:
    frames: Any
    orig_frames: Any = None
    cache_frames: Any = None


class VideoProcessor:
    """The subset of diffusers.video_processor.VideoProcessor the path uses: postprocess_video (denormalise to [0, 1],
    channels-last frames per batch item).  "pil" needs PIL; "np" / "pt" / "latent" do not."""

    def __init__(self, vae_scale_factor: int = 8):
        self.vae_scale_factor = vae_scale_factor

    def postprocess_video(self, video: torch.Tensor, output_type: str = "np"):
        if output_type == "latent":
            return video
        if output_type == "uint8":
            # schema extension: [B, F, H, W, 3] uint8 numpy frames packed on the GPU — bit-identical to
            # `(postprocess_video(..., "np") * 255).round().astype(uint8)` (what export_to_video does with "np" frames)
            return np.stack([E.vae_frames_to_rgb8(v.to(torch.bfloat16)).cpu().numpy() for v in video])
        v = (video.float() / 2 + 0.5).clamp(0, 1)          # [B, C, F, H, W]
        if output_type == "pt":
            return v.permute(0, 2, 1, 3, 4)
        arr = v.permute(0, 2, 3, 4, 1).cpu().numpy()         # [B, F, H, W, C]
        if output_type == "np":
            return arr
        if output_type == "pil":
            from PIL import Image
            return [[Image.fromarray((f * 255).round().astype("uint8")) for f in b] for b in arr]
        raise ValueError(f"unknown output_type {output_type}")


def get_resize_crop_region_for_grid(src, tgt_width, tgt_height):
    """pipeline_cogvideox_mp_fifo.py:84-100 (HunyuanDiT helper): centre crop region of `src`=(h, w) inside the base grid."""
    tw, th = tgt_width, tgt_height
    h, w = src
    r = h / w
    if r > (th / tw):
        resize_height, resize_width = th, int(round(th / h * w))
    else:
        resize_width, resize_height = tw, int(round(tw / w * h))
    crop_top = int(round((th - resize_height) / 2.0))
    crop_left = int(round((tw - resize_width) / 2.0))
    return (crop_top, crop_left), (crop_top + resize_height, crop_left + resize_width)


def retrieve_timesteps(scheduler, num_inference_steps=None, device=None, timesteps=None):
    if timesteps is not None:
        raise NotImplementedError("custom timestep lists are not used by the shipped configs")
    scheduler.set_timesteps(num_inference_steps, device=device)
    return scheduler.timesteps, num_inference_steps


class MPFIFOVideoIPAdapterCogVideoXPipeline:
    def __init__(self, tokenizer, text_encoder, vae, transformer, scheduler, resampler=None, image_encoder=None,
                 feature_extractor=None):
        if image_encoder is not None:
            raise NotImplementedError("the CLIP/DINO image-encoder branch is not on the shipped path (use_vae_as_encoder: true)")
        self.tokenizer, self.text_encoder, self.vae, self.transformer, self.scheduler = \
            tokenizer, text_encoder, vae, transformer, scheduler
        self.resampler, self.image_encoder, self.feature_extractor = resampler, image_encoder, feature_extractor
        self.vae_scale_factor_spatial = 2 ** (len(vae.config.block_out_channels) - 1) if vae is not None else 8
        self.vae_scale_factor_temporal = vae.config.temporal_compression_ratio if vae is not None else 4
        self.vae_scaling_factor_image = vae.config.scaling_factor if vae is not None else 0.7
        self.video_processor = VideoProcessor(vae_scale_factor=self.vae_scale_factor_spatial)
        self._device = transformer.device if transformer is not None else torch.device("cpu")
        self._guidance_scale, self._attention_kwargs, self._interrupt, self._num_timesteps = 6.0, None, False, 0
        # The reference samples the VAE posterior of the conditioning clips WITHOUT a generator (`latent_dist.sample()`,
        # pipeline_cogvideox_mp_fifo.py:585): the device's global RNG, not the call's `generator`, whose stream is consumed
        # only by prepare_latents and the scheduler steps.  None keeps exactly that; a generator here makes the condensed
        # tokens reproducible (schema extension; the parity test replays the reference's CPU global stream through it).
        self.vae_posterior_generator = None

    # ------------------------------------------------------------------ construction / placement
    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, transformer=None, resampler=None, torch_dtype=torch.bfloat16,
                        vae=None, text_encoder=None, tokenizer=None, scheduler=None, **kwargs):
        """infer_cogvideo_mp_fifo.py:171-176.  Loads `vae/` (mirror), `scheduler/scheduler_config.json`, and — when the
        `transformers` package and the files are present — `tokenizer/` + `text_encoder/` (third-party T5)."""
        from .loading import load_config
        from .transformer import CogVideoXTransformer3DModel
        from .vae import AutoencoderKLCogVideoX
        root = pretrained_model_name_or_path
        if transformer is None:
            transformer = CogVideoXTransformer3DModel.from_pretrained(root, subfolder="transformer", torch_dtype=torch_dtype,
                                                                      device=kwargs.get("device"))
        if vae is False:       # LongVGenCogVideoXPipeline (T2To): no VAE in that stage
            vae = None
        elif vae is None:
            vae = AutoencoderKLCogVideoX.from_pretrained(root, subfolder="vae", torch_dtype=torch_dtype, device=kwargs.get("device"))
        if scheduler is None:
            path = os.path.join(root, "scheduler", "scheduler_config.json")
            if not os.path.exists(path):
                raise FileNotFoundError(f"{path} is missing: the noise schedule comes from the checkpoint tree, there is no default")
            scheduler = CogVideoXDPMScheduler.from_config(load_config(root, "scheduler", "scheduler_config.json"),
                                                          timestep_spacing="trailing")
        if text_encoder is None and os.path.isdir(os.path.join(root, "text_encoder")):
            try:
                from transformers import T5EncoderModel, T5Tokenizer
                tokenizer = T5Tokenizer.from_pretrained(root, subfolder="tokenizer")
                text_encoder = T5EncoderModel.from_pretrained(root, subfolder="text_encoder", torch_dtype=torch_dtype)
            except Exception as e:  # third-party encoder is optional: prompt_embeds can be passed instead
                print(f"tokensgen_b200: T5 text encoder not loaded ({e}); pass prompt_embeds")
        if vae is None:
            return cls(tokenizer, text_encoder, transformer, scheduler)
        return cls(tokenizer, text_encoder, vae, transformer, scheduler, resampler=resampler)

    def to(self, device):
        self._device = torch.device(device)
        for m in (self.vae, self.transformer, self.resampler, self.text_encoder):
            if m is not None:
                m.to(self._device)
        return self

    @property
    def device(self):
        return self._device

    _execution_device = device

    @property
    def guidance_scale(self):
        return self._guidance_scale

    @property
    def num_timesteps(self):
        return self._num_timesteps

    @property
    def attention_kwargs(self):
        return self._attention_kwargs

    @property
    def interrupt(self):
        return self._interrupt

    # ------------------------------------------------------------------ prompt (:365-486)
    def _get_t5_prompt_embeds(self, prompt, num_videos_per_prompt=1, max_sequence_length=226, device=None, dtype=None):
        if self.text_encoder is None or self.tokenizer is None:
            raise E.TokensGenError("no text encoder in this pipeline: pass prompt_embeds / negative_prompt_embeds")
        device = device or self.device
        dtype = dtype or self.text_encoder.dtype
        prompt = [prompt] if isinstance(prompt, str) else prompt
        ids = self.tokenizer(prompt, padding="max_length", max_length=max_sequence_length, truncation=True,
                             add_special_tokens=True, return_tensors="pt").input_ids
        emb = self.text_encoder(ids.to(device))[0].to(dtype=dtype, device=device)
        _, seq_len, _ = emb.shape
        return emb.repeat(1, num_videos_per_prompt, 1).view(len(prompt) * num_videos_per_prompt, seq_len, -1)

    def encode_prompt(self, prompt, negative_prompt=None, do_classifier_free_guidance=True, num_videos_per_prompt=1,
                      prompt_embeds=None, negative_prompt_embeds=None, max_sequence_length=226, device=None, dtype=None):
        device = device or self.device
        batch = len(prompt) if isinstance(prompt, list) else (1 if prompt is not None else prompt_embeds.shape[0])
        if prompt_embeds is None:
            prompt_embeds = self._get_t5_prompt_embeds(prompt, num_videos_per_prompt, max_sequence_length, device, dtype)
        if do_classifier_free_guidance and negative_prompt_embeds is None:
            neg = negative_prompt or ""
            neg = batch * [neg] if isinstance(neg, str) else neg
            negative_prompt_embeds = self._get_t5_prompt_embeds(neg, num_videos_per_prompt, max_sequence_length, device, dtype)
        return prompt_embeds.to(device), (None if negative_prompt_embeds is None else negative_prompt_embeds.to(device))

    # ------------------------------------------------------------------ conditioning video -> condensed tokens (:562-648)
    def _encode_video_chunks(self, video, nf_per_chunk, device, dtype, generator):
        """pipeline_cogvideox_mp_fifo.py:576-588: pad one chunk, VAE-encode chunk by chunk, sample the posterior, scale."""
        video = video.to(device=device, dtype=dtype).permute(0, 2, 1, 3, 4)       # b c f h w
        video = torch.cat([video] + [video[:, :, [-1]]] * nf_per_chunk, dim=2)    # pad one chunk (:580-581)
        n = video.shape[2] // nf_per_chunk
        grp = getattr(self, "_clip_parallel_group", None)
        if grp is None:
            lat = []
            for c in range(n):
                dist = self.vae.encode(video[:, :, c * nf_per_chunk:(c + 1) * nf_per_chunk].contiguous()).latent_dist
                lat.append(dist.sample(generator=generator, scale=self.vae.config.scaling_factor))   # K19: sample * scaling fused
            return torch.cat(lat, dim=2).permute(0, 2, 1, 3, 4)                   # b f c h w
        # clip-parallel: every rank of the group was called with the same video; the chunks are independent (the conv
        # cache is cleared per encode call), so chunk c is encoded by group rank c % P and broadcast.  The posterior
        # noise of the other ranks' chunks is still drawn (and dropped) so the generator stream — and with it every
        # later draw — is the one the serial loop produces: results are identical to the single-process call.
        import torch.distributed as tdist
        P, r = tdist.get_world_size(grp), tdist.get_rank(grp)
        cfg = self.vae.config
        lshape = (video.shape[0], cfg.latent_channels, (nf_per_chunk - 1) // cfg.temporal_compression_ratio + 1,
                  video.shape[3] // self.vae_scale_factor_spatial, video.shape[4] // self.vae_scale_factor_spatial)
        lat = []
        for c in range(n):
            if c % P == r:
                dist = self.vae.encode(video[:, :, c * nf_per_chunk:(c + 1) * nf_per_chunk].contiguous()).latent_dist
                lat.append(dist.sample(generator=generator, scale=cfg.scaling_factor).contiguous())
                if tuple(lat[-1].shape) != lshape:
                    raise E.TokensGenError(f"clip-parallel encode: latent shape {tuple(lat[-1].shape)} != {lshape}")
            else:
                E.randn_tensor(lshape, generator, device, dtype)
                lat.append(torch.empty(lshape, device=device, dtype=dtype))
        for c in range(n):
            tdist.broadcast(lat[c], src=tdist.get_global_rank(grp, c % P), group=grp)
        return torch.cat(lat, dim=2).permute(0, 2, 1, 3, 4)

    def vae_encode_image(self, frames, device, do_classifier_free_guidance, use_separate_guidance, nf_per_chunk,
                         compressed_nf_per_chunk, num_chunks, resampler_image_rotary_emb, resampler_sampling_rotary_emb,
                         image_embeddings=None, generator=None):
        dtype = torch.bfloat16
        pe = self.transformer.patch_embed

        def encode_video(video):
            return self._encode_video_chunks(video, nf_per_chunk, device, dtype, self.vae_posterior_generator)

        def condense(lat):
            b, f, c, h, w = lat.shape
            p = pe.patch_size
            rows = E.patchify(lat.contiguous(), p)                                    # patch_embed.proj as a GEMM (:596)
            pw = pe.proj.weight.reshape(pe.proj.weight.shape[0], -1)
            tok = E.gemm_bias_act(rows, pw, pe.proj.bias).view(b, f, (h // p) * (w // p), -1)
            out = [self.resampler(tok[:, c0:c0 + compressed_nf_per_chunk], image_rotary_emb=resampler_image_rotary_emb,
                                  sampling_rotary_emb=resampler_sampling_rotary_emb)
                   for c0 in range(0, (f // compressed_nf_per_chunk) * compressed_nf_per_chunk, compressed_nf_per_chunk)]
            return torch.cat(out, dim=1)

        if image_embeddings is None:
            image_embeddings = condense(encode_video(frames))
        else:
            image_embeddings = image_embeddings.to(device=device, dtype=dtype)
            image_embeddings = torch.cat([image_embeddings] + [image_embeddings[:, [-1]]] *
                                         (image_embeddings.shape[1] // num_chunks), dim=1)          # :611-616
        if do_classifier_free_guidance:
            nt = self.resampler.config.num_temporal_queries
            if use_separate_guidance:
                zeros = torch.zeros((image_embeddings.shape[0], nf_per_chunk * num_chunks, 3, frames.shape[-2] if frames is not None else 480,
                                     frames.shape[-1] if frames is not None else 720))
                uncond = condense(encode_video(zeros))
                n_uncond = uncond.shape[1]
            else:
                uncond, n_uncond = None, (num_chunks + 1) * nt      # the reference only uses its length (:636-646)
            if image_embeddings.shape[1] != n_uncond:
                image_embeddings = torch.cat([image_embeddings] + [image_embeddings[:, [-1]]] *
                                             (n_uncond - image_embeddings.shape[1]), dim=1)
            if use_separate_guidance:
                image_embeddings = torch.cat([image_embeddings, uncond, image_embeddings], dim=0)
            else:
                image_embeddings = torch.cat([image_embeddings, image_embeddings], dim=0)
        return image_embeddings

    # ------------------------------------------------------------------ latents / decode / RoPE (:650-813)
    def prepare_latents(self, batch_size, num_channels_latents, num_chunks, num_frames_per_chunk, height, width, dtype, device,
                        generator, latents=None):
        shape = (batch_size, num_chunks * ((num_frames_per_chunk - 1) // self.vae_scale_factor_temporal + 1),
                 num_channels_latents, height // self.vae_scale_factor_spatial, width // self.vae_scale_factor_spatial)
        if latents is None:
            latents = E.randn_tensor(shape, generator, device, dtype)
        else:
            latents = latents.to(device)
        return latents * self.scheduler.init_noise_sigma

    def decode_latents(self, latents: torch.Tensor, nf_per_chunk=13) -> torch.Tensor:
        latents = latents.permute(0, 2, 1, 3, 4)
        latents = 1 / self.vae_scaling_factor_image * latents
        frames = [self.vae.decode(latents[:, :, c * nf_per_chunk:(c + 1) * nf_per_chunk].contiguous()).sample
                  for c in range(latents.shape[2] // nf_per_chunk)]
        return torch.cat(frames, dim=2)

    def _prepare_rotary_positional_embeddings(self, height, width, num_frames, device):
        p = self.transformer.config.patch_size
        gh, gw = height // (self.vae_scale_factor_spatial * p), width // (self.vae_scale_factor_spatial * p)
        base_w, base_h = 720 // (self.vae_scale_factor_spatial * p), 480 // (self.vae_scale_factor_spatial * p)
        (top, left), (bottom, right) = get_resize_crop_region_for_grid((gh, gw), base_w, base_h)
        return get_3d_rotary_pos_embed(self.transformer.config.attention_head_dim,
                                       [[0, top, left], [num_frames, bottom, right]], (num_frames, gh, gw), device=device)

    def _prepare_vip_rotary_positional_embeddings(self, grid_t, grid_h, grid_w, device):
        return get_3d_rotary_pos_embed_v2(self.transformer.config.attention_head_dim, grid_t, grid_h, grid_w, device=device)

    def _vip_grids(self, latents, num_chunks, compressed_nf_per_chunk, start_idx):
        """:1061-1149 -> the two position grids and the two Resampler RoPE tables."""
        p = self.transformer.config.patch_size
        rc = self.resampler.config
        lin = lambda a, b, n: np.linspace(a, b, n, endpoint=False, dtype=np.float32)
        gh, gw = latents.shape[-2] // p, latents.shape[-1] // p
        img = [lin(0, num_chunks * compressed_nf_per_chunk, num_chunks * compressed_nf_per_chunk), lin(0, gh, gh), lin(0, gw, gw)]
        cond = [np.concatenate([lin(start_idx + i * compressed_nf_per_chunk, start_idx + (i + 1) * compressed_nf_per_chunk,
                                    rc.num_temporal_queries) for i in range(num_chunks + 1)]),
                lin(0, gh, rc.num_height_queries), lin(0, gw, rc.num_width_queries)]
        dev = self.device
        rs_img = self._prepare_vip_rotary_positional_embeddings(lin(0, rc.max_temporal_seq_len, rc.max_temporal_seq_len),
                                                                lin(0, rc.max_height_seq_len, rc.max_height_seq_len),
                                                                lin(0, rc.max_width_seq_len, rc.max_width_seq_len), dev)
        rs_smp = self._prepare_vip_rotary_positional_embeddings(
            lin(start_idx, start_idx + rc.max_temporal_seq_len, rc.num_temporal_queries),
            lin(0, rc.max_height_seq_len, rc.num_height_queries), lin(0, rc.max_width_seq_len, rc.num_width_queries), dev)
        return img, cond, rs_img, rs_smp

    def _set_vip_scale(self, vip_scale):
        for _, module in self.transformer.named_modules():               # :981-983: matched by class NAME
            if module.__class__.__name__ == "VideoIPAdapterCogVideoXAttnProcessor2_0":
                module.scale = vip_scale

    # ------------------------------------------------------------------ base stage (:837-1344)
    @torch.no_grad()
    def __call__(self, prompt=None, frames=None, negative_prompt=None, height: int = 480, width: int = 720,
                 num_frames_per_chunk: int = 49, max_num_chunks: Optional[int] = 1, max_num_chunks_w_fifo: Optional[int] = None,
                 max_num_chunks_wo_fifo: Optional[int] = 1, num_inference_steps: int = 50, timesteps=None,
                 use_separate_guidance: bool = False, guidance_scale: float = 6, guidance_scale_img: float = 6,
                 use_dynamic_cfg: bool = False, num_videos_per_prompt: int = 1, eta: float = 0.0, generator=None,
                 latents=None, prompt_embeds=None, negative_prompt_embeds=None, image_embeddings=None,
                 output_type: str = "pil", return_dict: bool = True, attention_kwargs=None, callback_on_step_end=None,
                 callback_on_step_end_tensor_inputs=("latents",), max_sequence_length: int = 226, decode_chunk_size=None,
                 vip_scale=1.0, sampling_mode: str = None, sampling_params: Dict[str, Any] = None, cache_idx=(),
                 video_ipadapter_start_frame_idx: Optional[int] = 1000, cfg_parallel_group=None,
                 sequence_parallel_group=None):
        if callback_on_step_end is not None:
            raise NotImplementedError("step callbacks are not used on the reproduced path")
        nf_per_chunk = num_frames_per_chunk
        num_chunks = frames.shape[1] // nf_per_chunk if frames is not None else max_num_chunks
        num_chunks_wo_fifo = max(1, min(max_num_chunks_wo_fifo, num_chunks))
        num_chunks_w_fifo = max(1, min(max_num_chunks_w_fifo, num_chunks)) if max_num_chunks_w_fifo is not None else num_chunks
        if sampling_mode is not None and "freeinit" in sampling_mode:
            raise NotImplementedError("freeinit sampling is not on the shipped path (sampling_mode: fifo)")
        if num_chunks_wo_fifo != 1:
            raise NotImplementedError("the shipped configs prime the FIFO from ONE base clip (max_num_chunks_wo_fifo: 1)")
        cache_idx = [] if cache_idx is None else list(cache_idx)
        compressed_nf_per_chunk = (nf_per_chunk - 1) // self.vae_scale_factor_temporal + 1
        use_vip = self.resampler is not None
        if num_frames_per_chunk > 49:
            raise ValueError("The number of frames must be less than 49 for now due to static positional embeddings.")
        if height % 8 != 0 or width % 8 != 0:
            raise ValueError(f"`height` and `width` have to be divisible by 8 but are {height} and {width}.")
        if prompt is None and prompt_embeds is None:
            raise ValueError("Provide either `prompt` or `prompt_embeds`.")
        self._set_vip_scale(vip_scale)
        self._guidance_scale, self._attention_kwargs, self._interrupt = guidance_scale, attention_kwargs, False
        device = self.device
        do_cfg = guidance_scale > 1.0
        batch_size = 1 if isinstance(prompt, str) else (len(prompt) if prompt is not None else prompt_embeds.shape[0])

        prompt_embeds, negative_prompt_embeds = self.encode_prompt(
            prompt, negative_prompt, do_cfg, num_videos_per_prompt=1, prompt_embeds=prompt_embeds,
            negative_prompt_embeds=negative_prompt_embeds, max_sequence_length=max_sequence_length, device=device)
        use_separate_guidance = bool(use_separate_guidance) and do_cfg
        if do_cfg:
            if use_separate_guidance:   # :1026-1027 — branches (uncond_txt, uncond_img, txt_img)
                prompt_embeds = torch.cat([negative_prompt_embeds, prompt_embeds, prompt_embeds], dim=0)
            else:
                prompt_embeds = torch.cat([negative_prompt_embeds, prompt_embeds], dim=0)
        prompt_embeds = prompt_embeds.to(torch.bfloat16)

        timesteps, num_inference_steps = retrieve_timesteps(self.scheduler, num_inference_steps, device, timesteps)
        self._num_timesteps = len(timesteps)
        latents = self.prepare_latents(batch_size, self.transformer.config.in_channels, num_chunks_wo_fifo, nf_per_chunk, height,
                                       width, prompt_embeds.dtype, device, generator, latents)
        image_rotary_emb = self._prepare_rotary_positional_embeddings(height, width, compressed_nf_per_chunk, device)

        vip_nf_per_chunk = None
        img_grid = cond_grid = None
        if use_vip:
            img_grid, cond_grid, rs_img, rs_smp = self._vip_grids(latents, num_chunks, compressed_nf_per_chunk,
                                                                  video_ipadapter_start_frame_idx)
            vip_nf_per_chunk = self.resampler.config.num_temporal_queries
            self._clip_parallel_group = sequence_parallel_group   # conditioning chunks are encoded one per rank (vae_encode_image)
            try:
                image_embeddings = self.vae_encode_image(frames, device, do_cfg, use_separate_guidance, nf_per_chunk,
                                                         compressed_nf_per_chunk, num_chunks, rs_img, rs_smp,
                                                         image_embeddings=image_embeddings, generator=generator)
            finally:
                self._clip_parallel_group = None
            n_vip_frames = min(vip_nf_per_chunk + 1, compressed_nf_per_chunk)
            cond_rope = self._prepare_vip_rotary_positional_embeddings(cond_grid[0][:n_vip_frames], cond_grid[1], cond_grid[2], device)
            img_rope = self._prepare_vip_rotary_positional_embeddings(img_grid[0][:compressed_nf_per_chunk], img_grid[1], img_grid[2], device)
            vip_states = image_embeddings[:, :n_vip_frames].contiguous()

        # denoising loop with the diagonal FIFO capture (:1186-1305)
        fifo_latents: List[torch.Tensor] = []
        fifo_old: List[Optional[torch.Tensor]] = []
        old_x0 = None
        ts = [int(t) for t in timesteps]
        B = (3 if use_separate_guidance else 2) if do_cfg else 1
        if use_separate_guidance and cfg_parallel_group is not None:
            cfg_parallel_group = None   # the two-rank split is for the two-branch bundle; three branches run on one rank
        # CFG-parallel base stage: the two guidance branches are independent until they are combined, so with
        # `cfg_parallel_group` (two ranks that call this method with identical arguments) each rank runs ONE branch through the
        # DiT (B = 1) and the pair all-gathers the 2.25 MB predictions over NVLink; everything else (latents, scheduler
        # noise, FIFO capture) is computed redundantly and identically on both, so no other state moves.
        cfgp = cfg_parallel_group if (do_cfg and cfg_parallel_group is not None) else None
        # Sequence-parallel base stage (tokensgen_b200/seqpar.py): every rank of `sequence_parallel_group`
        # calls this method with identical arguments; each DiT forward is sharded over the group (rows for LayerNorm /
        # GEMMs, heads for attention, the two all-to-alls fused into kernel epilogues as NVLink peer stores) and returns
        # the full, bit-identical prediction on every rank.  Latents, noise and the FIFO capture are computed redundantly.
        if sequence_parallel_group is not None:
            if cfgp is not None:
                raise ValueError("choose one of cfg_parallel_group / sequence_parallel_group for the base stage")
            self.transformer.enable_sequence_parallel(sequence_parallel_group)
        if cfgp is not None:
            import torch.distributed as dist
            branch = dist.get_rank(cfgp)
            if dist.get_world_size(cfgp) != 2:
                raise ValueError("cfg_parallel_group must hold exactly two ranks (uncond, cond)")
            pe_mine = prompt_embeds[branch:branch + 1].contiguous()
            vip_mine = vip_states[branch:branch + 1].contiguous() if use_vip else None
        for i, t in enumerate(ts):
            k = max(0, compressed_nf_per_chunk - 1 - i)
            fifo_latents.insert(0, latents[:, [k]])
            fifo_old.insert(0, None if old_x0 is None else old_x0[:, [k]])
            if cfgp is not None:
                kw = dict(vip_image_rotary_emb=img_rope, vip_condition_rotary_emb=cond_rope,
                          vip_encoder_hidden_states=vip_mine) if use_vip else {}
                mine = self.transformer(hidden_states=latents, encoder_hidden_states=pe_mine,
                                        timestep=torch.full((1,), t, device=device, dtype=torch.int64),
                                        image_rotary_emb=image_rotary_emb, attention_kwargs=attention_kwargs,
                                        return_dict=False, **kw)[0].contiguous()
                both = torch.empty((2,) + tuple(mine.shape[1:]), device=device, dtype=mine.dtype)
                dist.all_gather_into_tensor(both, mine, group=cfgp)
                noise_pred = both.float()
            else:
                model_in = torch.cat([latents] * B) if do_cfg else latents
                timestep = torch.full((B,), t, device=device, dtype=torch.int64)
                kw = dict(vip_image_rotary_emb=img_rope, vip_condition_rotary_emb=cond_rope,
                          vip_encoder_hidden_states=vip_states) if use_vip else {}
                noise_pred = self.transformer(hidden_states=model_in, encoder_hidden_states=prompt_embeds, timestep=timestep,
                                              image_rotary_emb=image_rotary_emb, attention_kwargs=attention_kwargs,
                                              return_dict=False, **kw)[0]
                noise_pred = noise_pred.float()
            g, g_img = guidance_scale, guidance_scale_img
            if use_dynamic_cfg:
                ramp = (1 - math.cos(math.pi * ((num_inference_steps - t) / num_inference_steps) ** 5.0)) / 2
                g, g_img = 1 + guidance_scale * ramp, 1 + guidance_scale_img * ramp
                self._guidance_scale = g
            if do_cfg and use_separate_guidance:     # :1261-1263
                ut, ui, a = noise_pred.chunk(3)
                noise_pred = a + (g - 1) * (a - ut) + (g_img - 1) * (a - ui)
            elif do_cfg:
                u, c = noise_pred.chunk(2)
                noise_pred = u + g * (c - u)
            prev_t = ts[i + 1] if i + 1 < len(ts) else -1
            latents, old_x0 = self.scheduler.step(noise_pred, old_x0, t, prev_t, ts[i - 1] if i > 0 else None, latents,
                                                  generator=generator, return_dict=False)
            latents = latents.to(prompt_embeds.dtype)
        if sequence_parallel_group is not None:
            self.transformer.disable_sequence_parallel()   # the FIFO stage is window-parallel: one whole window per rank
        orig_latents = latents.clone()

        return FIFOCogVideoXPipelineOutput(
            fifo_latents=torch.cat(fifo_latents, dim=1), fifo_old_pred_original_sample=fifo_old, orig_latents=orig_latents,
            nf_per_chunk=compressed_nf_per_chunk, vip_nf_per_chunk=vip_nf_per_chunk,
            num_frames=num_chunks_w_fifo * compressed_nf_per_chunk, image_embeddings=image_embeddings if use_vip else None,
            timesteps=timesteps, num_inference_steps=num_inference_steps, do_classifier_free_guidance=do_cfg,
            use_separate_guidance=use_separate_guidance, use_dynamic_cfg=use_dynamic_cfg, prompt_embeds=prompt_embeds,
            image_rotary_emb=image_rotary_emb, vip_image_rotary_grid=img_grid, vip_condition_rotary_grid=cond_grid,
            cache_idx=cache_idx, attention_kwargs=attention_kwargs, guidance_scale=guidance_scale,
            guidance_scale_img=guidance_scale_img, extra_step_kwargs={"generator": generator}, condition_frames=frames,
            video_ipadapter_start_frame_idx=video_ipadapter_start_frame_idx, sampling_params=sampling_params,
            output_type=output_type, return_dict=return_dict)

    @torch.no_grad()
    def preprare_for_fifo(self, prompt=None, frames=None, num_inference_steps: int = 50, timesteps=None, guidance_scale: float = 6,
                          attention_kwargs=None, vip_scale=1.0, **_unused):
        """(sic) :1348-1518 — what the non-primary pipelines do before the FIFO stage: vip scale + scheduler timesteps."""
        self._set_vip_scale(vip_scale)
        self._guidance_scale, self._attention_kwargs, self._interrupt = guidance_scale, attention_kwargs, False
        timesteps, _ = retrieve_timesteps(self.scheduler, num_inference_steps, self.device, timesteps)
        self._num_timesteps = len(timesteps)
