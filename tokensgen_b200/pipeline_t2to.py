"""Host-side mirror of the reference's T2To pipeline, `LongVGenCogVideoXPipeline`
(longvgen/pipeline/pipeline_cogvideox_t2to.py:584-912): text -> a [1, 96, 16, 8, 12] grid of PCA-space condensed tokens,
denoised by the same DiT architecture configured with patch_size = 1 (9 216 tokens + 226 text, plain attention
processor, RoPE dims 52/6/6 on the t/h/w axes, :543-564), dynamic classifier-free guidance (:849-858), then
un-normalisation and the PCA inverse back to 3072-d (:891-904).  The DiT forward and the scheduler step are the CUDA
mirrors; the tail is three tiny fp32 host ops exactly as in the reference (it runs them on the CPU)."""
from __future__ import annotations

import math
from typing import Optional

import numpy as np
import torch

from . import _ext as E
from .pipeline import CogVideoXPipelineOutput, MPFIFOVideoIPAdapterCogVideoXPipeline, retrieve_timesteps
from .rope import get_3d_rotary_pos_embed_v2


class LongVGenCogVideoXPipeline(MPFIFOVideoIPAdapterCogVideoXPipeline):
    def __init__(self, tokenizer, text_encoder, transformer, scheduler, **_unused):
        """Same positional signature as the reference (pipeline_cogvideox_t2to.py:297-311): this stage has no VAE."""
        super().__init__(tokenizer, text_encoder, None, transformer, scheduler, resampler=None)

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, transformer=None, torch_dtype=torch.bfloat16, text_encoder=None,
                        tokenizer=None, scheduler=None, **kwargs):
        """infer_cogvideo_mp_fifo.py:225-229: the T2To stage shares the checkpoint tree's scheduler / T5 with the To2V one."""
        kwargs.pop("vae", None), kwargs.pop("resampler", None)
        return super().from_pretrained(pretrained_model_name_or_path, transformer=transformer, torch_dtype=torch_dtype,
                                       vae=False, text_encoder=text_encoder, tokenizer=tokenizer, scheduler=scheduler, **kwargs)

    def prepare_latents(self, batch_size, num_channels_latents, num_chunks, num_frames_per_chunk, height, width, dtype, device,
                        generator, latents=None):
        shape = (batch_size, num_chunks * num_frames_per_chunk, num_channels_latents, height, width)      # :445-451
        latents = E.randn_tensor(shape, generator, device, dtype) if latents is None else latents.to(device)
        return latents * self.scheduler.init_noise_sigma

    def _prepare_rotary_positional_embeddings(self, grid_t, grid_h, grid_w, device):
        return get_3d_rotary_pos_embed_v2(self.transformer.config.attention_head_dim, grid_t, grid_h, grid_w,
                                          dim_t=52, dim_h=6, dim_w=6, device=device)

    @torch.no_grad()
    def __call__(self, prompt=None, negative_prompt=None, height: int = 480, width: int = 720, num_frames_per_chunk: int = 49,
                 num_chunks: Optional[int] = 1, num_inference_steps: int = 50, timesteps=None, guidance_scale: float = 6,
                 use_dynamic_cfg: bool = False, num_videos_per_prompt: int = 1, eta: float = 0.0, generator=None, latents=None,
                 prompt_embeds=None, negative_prompt_embeds=None, return_dict: bool = True, attention_kwargs=None,
                 callback_on_step_end=None, callback_on_step_end_tensor_inputs=("latents",), max_sequence_length: int = 226,
                 longvgen_mean=None, longvgen_std=None, longvgen_pca=None, sequence_parallel_group=None):
        if num_frames_per_chunk > 4:
            raise ValueError("The number of frames must equal 4 for now due to static positional embeddings.")
        if callback_on_step_end is not None:
            raise NotImplementedError("step callbacks are not used on the reproduced path")
        load = lambda x, **kw: torch.load(x, **kw) if isinstance(x, (str, bytes)) or hasattr(x, "__fspath__") else x
        mean, std = load(longvgen_mean, weights_only=True), load(longvgen_std, weights_only=True)
        pca = load(longvgen_pca, weights_only=False)
        num_frames = num_chunks * num_frames_per_chunk
        device = self.device
        self._guidance_scale, self._attention_kwargs, self._interrupt = guidance_scale, attention_kwargs, False
        do_cfg = guidance_scale > 1.0
        batch_size = 1 if isinstance(prompt, str) else (len(prompt) if prompt is not None else prompt_embeds.shape[0])
        prompt_embeds, negative_prompt_embeds = self.encode_prompt(
            prompt, negative_prompt, do_cfg, num_videos_per_prompt=1, prompt_embeds=prompt_embeds,
            negative_prompt_embeds=negative_prompt_embeds, max_sequence_length=max_sequence_length, device=device)
        if do_cfg:
            prompt_embeds = torch.cat([negative_prompt_embeds, prompt_embeds], dim=0)
        prompt_embeds = prompt_embeds.to(torch.bfloat16)
        timesteps, num_inference_steps = retrieve_timesteps(self.scheduler, num_inference_steps, device, timesteps)
        self._num_timesteps = len(timesteps)
        latents = self.prepare_latents(batch_size, 16, num_chunks, num_frames_per_chunk, height, width, prompt_embeds.dtype,
                                       device, generator, latents)
        lin = lambda n: np.linspace(0, n, n, endpoint=False, dtype=np.float32)
        rope = self._prepare_rotary_positional_embeddings(lin(num_frames), lin(latents.shape[-2]), lin(latents.shape[-1]), device)
        ts = [int(t) for t in timesteps]
        old_x0 = None
        B = 2 if do_cfg else 1
        if sequence_parallel_group is not None:   # every rank of the group calls with identical arguments (seqpar.py)
            self.transformer.enable_sequence_parallel(sequence_parallel_group)
        for i, t in enumerate(ts):
            model_in = torch.cat([latents] * 2) if do_cfg else latents
            noise_pred = self.transformer(hidden_states=model_in, encoder_hidden_states=prompt_embeds,
                                          timestep=torch.full((B,), t, device=device, dtype=torch.int64),
                                          image_rotary_emb=rope, attention_kwargs=attention_kwargs, return_dict=False)[0].float()
            g = guidance_scale
            if use_dynamic_cfg:
                g = 1 + guidance_scale * ((1 - math.cos(math.pi * ((num_inference_steps - t) / num_inference_steps) ** 5.0)) / 2)
                self._guidance_scale = g
            if do_cfg:
                u, c = noise_pred.chunk(2)
                noise_pred = u + g * (c - u)
            prev_t = ts[i + 1] if i + 1 < len(ts) else -1
            latents, old_x0 = self.scheduler.step(noise_pred, old_x0, t, prev_t, ts[i - 1] if i > 0 else None, latents,
                                                  generator=generator, return_dict=False)
            latents = latents.to(prompt_embeds.dtype)
        if sequence_parallel_group is not None:
            self.transformer.disable_sequence_parallel()
        # :891-904 — un-normalise the 16 PCA coordinates, zero-pad to the PCA width, inverse transform (fp32, host)
        dtype = latents.dtype
        b, f, c, h, w = latents.shape
        flat = latents.to(torch.float32).cpu().permute(0, 1, 3, 4, 2).reshape(-1, c)
        flat = flat * std[:, :16] + mean[:, :16]
        wide = torch.zeros((flat.shape[0], pca.components_.shape[0]), dtype=flat.dtype)  # 3072 in the shipped pca.pt
        wide[:, :16] = flat
        out = pca.inverse_transform(wide)
        out = out.reshape(b, f, h, w, -1).permute(0, 1, 4, 2, 3).to(device=device, dtype=dtype)
        if not return_dict:
            return (out,)
        return CogVideoXPipelineOutput(frames=out)
