"""CogVideoXDPMScheduler mirror (longvgen/schedulers/scheduling_dpm_cogvideox.py:136-541) on the C-ABI kernels.

Same constructor arguments, `set_timesteps`, `scale_model_input`, `step` argument order and return convention,
`add_noise_to_xt`, `init_noise_sigma`, `order`.  The scalar coefficient algebra stays on the host in fp64 exactly as the
reference computes it (0-dim fp64 tensors); the tensor arithmetic is one tg_cfg_dpm_step launch that reproduces the
reference's op-by-op rounding, so with the same noise the result is bit-identical.

`window_step` is the fused form the FIFO worker needs: classifier-free guidance + all 13 per-frame steps in one launch
(replaces cogvideo_sampling_mp_fifo.py:527-550).
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _ext as E


# scheduler/scheduler_config.json of THUDM/CogVideoX-5b (SURVEY §8 [HF config]); timestep_spacing is overridden to
# "trailing" by the CLI (infer_cogvideo_mp_fifo.py:177-178)
COGVIDEOX_5B_CONFIG = dict(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
                           clip_sample=False, set_alpha_to_one=True, steps_offset=0, prediction_type="v_prediction",
                           clip_sample_range=1.0, sample_max_value=1.0, timestep_spacing="trailing",
                           rescale_betas_zero_snr=True, snr_shift_scale=1.0)


class CogVideoXDPMScheduler:
    order = 1

    def __init__(self, num_train_timesteps: int = 1000, beta_start: float = 0.00085, beta_end: float = 0.0120,
                 beta_schedule: str = "scaled_linear", trained_betas=None, clip_sample: bool = True,
                 set_alpha_to_one: bool = True, steps_offset: int = 0, prediction_type: str = "epsilon",
                 clip_sample_range: float = 1.0, sample_max_value: float = 1.0, timestep_spacing: str = "leading",
                 rescale_betas_zero_snr: bool = False, snr_shift_scale: float = 3.0):
        """Defaults are the reference's own (scheduling_dpm_cogvideox.py:181-197), so a partial scheduler_config.json means
        the same schedule here as there; the CogVideoX-5b checkpoint's values are `COGVIDEOX_5B_CONFIG` / `.cogvideox_5b()`."""
        cfg = {k: v for k, v in locals().items() if k != "self"}
        self.config = SimpleNamespace(**cfg)
        if prediction_type != "v_prediction":
            raise NotImplementedError("the CogVideoX-5b path steps with prediction_type='v_prediction' (scheduler_config.json); "
                                      f"got {prediction_type!r} — pass the checkpoint's scheduler config or use .cogvideox_5b()")
        if trained_betas is not None:
            self.betas = torch.tensor(trained_betas, dtype=torch.float32)
        elif beta_schedule == "linear":
            self.betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
        elif beta_schedule == "scaled_linear":
            self.betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float64) ** 2
        else:
            raise NotImplementedError(f"{beta_schedule} is not implemented for {self.__class__}")
        self.alphas = 1.0 - self.betas
        ac = torch.cumprod(self.alphas, dim=0)
        ac = ac / (snr_shift_scale + (1 - snr_shift_scale) * ac)      # :217
        if rescale_betas_zero_snr:                                     # :95-122
            s = ac.sqrt()
            s0, sT = s[0].clone(), s[-1].clone()
            s = (s - sT) * (s0 / (s0 - sT))
            ac = s ** 2
        self.alphas_cumprod = ac
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.init_noise_sigma = 1.0
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].copy().astype(np.int64))

    @classmethod
    def cogvideox_5b(cls, **overrides):
        """The scheduler of the CogVideoX-5b checkpoint as the CLI configures it (trailing spacing)."""
        return cls(**{**COGVIDEOX_5B_CONFIG, **overrides})

    @classmethod
    def from_config(cls, config=None, **overrides):
        """`CogVideoXDPMScheduler.from_config(pipe.scheduler.config, timestep_spacing="trailing")`
        (infer_cogvideo_mp_fifo.py:177-178): a dict / namespace of constructor arguments.  Missing keys take the REFERENCE's
        defaults; keys this class does not know are dropped only if they are diffusers bookkeeping (`_class_name`, ...) —
        anything else is an error rather than a silently different schedule."""
        import inspect
        cfg = dict(vars(config)) if hasattr(config, "__dict__") and not isinstance(config, dict) else dict(config or {})
        cfg.update(overrides)
        names = set(inspect.signature(cls.__init__).parameters) - {"self"}
        unknown = [k for k in cfg if k not in names and not k.startswith("_")]
        if unknown:
            raise ValueError(f"CogVideoXDPMScheduler.from_config: unknown keys {unknown}")
        return cls(**{k: v for k, v in cfg.items() if k in names})

    # ------------------------------------------------------------------ reference API
    def scale_model_input(self, sample: torch.Tensor, timestep=None) -> torch.Tensor:
        return sample

    def set_timesteps(self, num_inference_steps: int, device=None):
        n_train = self.config.num_train_timesteps
        if num_inference_steps > n_train:
            raise ValueError(f"`num_inference_steps`: {num_inference_steps} cannot be larger than {n_train}")
        self.num_inference_steps = num_inference_steps
        sp = self.config.timestep_spacing
        if sp == "linspace":
            ts = np.linspace(0, n_train - 1, num_inference_steps).round()[::-1].copy().astype(np.int64)
        elif sp == "leading":
            ts = (np.arange(0, num_inference_steps) * (n_train // num_inference_steps)).round()[::-1].copy().astype(np.int64)
            ts += self.config.steps_offset
        elif sp == "trailing":
            ts = np.round(np.arange(n_train, 0, -n_train / num_inference_steps)).astype(np.int64) - 1
        else:
            raise ValueError(f"{sp} is not supported. Please make sure to choose one of 'leading' or 'trailing'.")
        self.timesteps = torch.from_numpy(ts).to(device)

    def coefficients(self, timestep: int, prev_timestep: int, timestep_back: Optional[int]) -> List[float]:
        """get_variables + get_mult + the x0 scalars, fp64 (scheduling_dpm_cogvideox.py:334-356, 429-448)."""
        a_t = self.alphas_cumprod[int(timestep)]
        a_prev = self.alphas_cumprod[int(prev_timestep)] if prev_timestep >= 0 else self.final_alpha_cumprod.double()
        lamb = ((a_t / (1 - a_t)) ** 0.5).log()
        lamb_next = ((a_prev / (1 - a_prev)) ** 0.5).log()
        h = lamb_next - lamb
        mult0 = ((1 - a_prev) / (1 - a_t)) ** 0.5 * (-h).exp()
        mult1 = (-2 * h).expm1() * a_prev ** 0.5
        mult2 = mult3 = torch.tensor(0.0, dtype=torch.float64)
        if timestep_back is not None:
            a_back = self.alphas_cumprod[int(timestep_back)]
            r = (lamb - ((a_back / (1 - a_back)) ** 0.5).log()) / h
            mult2, mult3 = 1 + 1 / (2 * r), 1 / (2 * r)
        mult_noise = (1 - a_prev) ** 0.5 * (1 - (-2 * h).exp()) ** 0.5
        return [float(v) for v in (a_t ** 0.5, (1 - a_t) ** 0.5, mult0, mult1, mult2, mult3, mult_noise)]

    def _coef_row(self, timestep, prev_timestep, timestep_back, has_old: bool) -> List[float]:
        second = has_old and prev_timestep >= 0
        if second and timestep_back is None:
            raise IndexError("second-order step without timestep_back (the reference fails the same way: mult[2])")
        return self.coefficients(timestep, prev_timestep, timestep_back) + [1.0 if second else 0.0]

    def step(self, model_output: torch.Tensor, old_pred_original_sample: Optional[torch.Tensor], timestep: int,
             prev_timestep: int, timestep_back: Optional[int], sample: torch.Tensor, eta: float = 0.0,
             use_clipped_model_output: bool = False, generator=None, variance_noise=None, return_dict: bool = False):
        if self.num_inference_steps is None:
            raise ValueError("Number of inference steps is 'None', you need to run 'set_timesteps' after creating the scheduler")
        if not sample.is_cuda:
            raise E.TokensGenError("CogVideoXDPMScheduler (tokensgen_b200) steps on CUDA only (no CPU fallback)")
        timestep, prev_timestep = int(timestep), int(prev_timestep)
        timestep_back = None if timestep_back is None else int(timestep_back)
        has_old = old_pred_original_sample is not None
        row = self._coef_row(timestep, prev_timestep, timestep_back, has_old)
        coef = torch.tensor([row], dtype=torch.float64).float().to(sample.device)
        # RNG consumption follows the reference: one draw always, a second one on the 2M branch (:450,:461)
        n1 = E.randn_tensor(sample.shape, generator, sample.device, sample.dtype)
        n2 = E.randn_tensor(sample.shape, generator, sample.device, sample.dtype) if row[7] else n1
        smp = sample.to(torch.bfloat16).reshape(1, -1).contiguous()
        flat = lambda t: t.reshape(1, -1).contiguous()
        if model_output.dtype == torch.bfloat16:
            prev, x0 = E.cfg_dpm_step(flat(model_output).unsqueeze(0), smp, flat(old_pred_original_sample) if row[7] else None,
                                      flat(n1.to(torch.bfloat16)), flat(n2.to(torch.bfloat16)), coef, 0.0, E.DPM_BF16_CHAIN)
        else:
            old = flat(old_pred_original_sample.float()) if row[7] else None
            prev, x0 = E.cfg_dpm_step(flat(model_output.float()).unsqueeze(0), smp, old, flat(n1.to(torch.bfloat16)),
                                      flat(n2.to(torch.bfloat16)), coef, 0.0, E.DPM_BASE_CHAIN)
            prev = prev.float()  # the reference's mixed chain yields fp32; the pipeline casts it back to bf16
        prev, x0 = prev.view(sample.shape), x0.view(sample.shape)
        if not return_dict:
            return (prev, x0)
        return SimpleNamespace(prev_sample=prev, pred_original_sample=x0)

    def window_step(self, noise_pred: torch.Tensor, latents: torch.Tensor, old_x0: Sequence[Optional[torch.Tensor]],
                    t: Sequence[int], prev_t: Sequence[int], next_t: Sequence[int], guidance_scale: float,
                    noise: Optional[Tuple[torch.Tensor, torch.Tensor]] = None, generator=None,
                    guidance_scale_img: Optional[float] = None):
        """noise_pred [2,F,C,H,W] (uncond, cond), [1,F,...], or [3,F,...] (uncond_txt, uncond_img, txt_img — the reference's
        use_separate_guidance, cogvideo_sampling_mp_fifo.py:528-530, which needs `guidance_scale_img`); latents [1,F,C,H,W]
        bf16; old_x0: F entries ([1,1,C,H,W] or None); t/prev_t/next_t: the window's rows of the FIFO timestep tables
        (next_t <= 0 means no history step).
        Returns (latents_out [1,F,...], [x0_j [1,1,...]] * F) exactly as the reference worker's loop does."""
        F = latents.shape[1]
        dev = latents.device
        rows = []
        for j in range(F):
            back = int(next_t[j]) if next_t[j] > 0 else None
            rows.append(self._coef_row(int(t[j]), int(prev_t[j]), back, old_x0[j] is not None))
        coef = torch.tensor(rows, dtype=torch.float64).float().to(dev)
        if noise is None:
            n1 = E.randn_tensor(latents.shape, generator, dev, latents.dtype)
            n2 = E.randn_tensor(latents.shape, generator, dev, latents.dtype)
        else:
            n1, n2 = noise
        frame_shape = latents.shape[2:]
        old = torch.zeros((F,) + tuple(frame_shape), device=dev, dtype=torch.bfloat16)
        for j in range(F):
            if rows[j][7]:
                old[j].copy_(old_x0[j].reshape(frame_shape))
        nb = noise_pred.shape[0]
        g1, g2 = float(guidance_scale), 0.0
        if nb == 3:
            if guidance_scale_img is None:
                raise ValueError("three guidance branches need guidance_scale_img")
            g1, g2 = float(guidance_scale) - 1.0, float(guidance_scale_img) - 1.0     # python doubles, like the reference
        prev, x0 = E.cfg_dpm_step(noise_pred.reshape(nb, F, -1).contiguous(), latents.reshape(F, -1).contiguous(),
                                  old.view(F, -1), n1.reshape(F, -1).contiguous(), n2.reshape(F, -1).contiguous(), coef,
                                  g1, E.DPM_BF16_CHAIN, guidance_scale2=g2)
        x0 = x0.view((F, 1, 1) + tuple(frame_shape))
        return prev.view(latents.shape), [x0[j] for j in range(F)]

    def add_noise_to_xt(self, xt_previous: torch.Tensor, noise: torch.Tensor, timesteps) -> torch.Tensor:
        """scheduling_dpm_cogvideox.py:497-518.  Returns the re-noised frame rounded once from the fp64 expression (the
        reference returns the fp64 tensor and its caller's in-place assignment performs that rounding)."""
        t = int(torch.as_tensor(timesteps).reshape(-1)[0])
        b = self.betas[t].double()
        q = xt_previous.to(torch.bfloat16).reshape(1, -1).contiguous().clone()
        E.queue_shift_renoise(q, None, noise.to(torch.bfloat16).reshape(-1).contiguous(), float((1 - b) ** 0.5), float(b ** 0.5))
        return q.view(xt_previous.shape)

    def __len__(self):
        return self.config.num_train_timesteps
