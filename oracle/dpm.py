"""ORACLE (test infrastructure): restatement of CogVideoXDPMScheduler for the FIFO path.

Follows longvgen/schedulers/scheduling_dpm_cogvideox.py: __init__ :181-260 (scaled_linear betas in fp64, SNR shift
:217, zero-terminal-SNR rescale :95-122/:220-221), set_timesteps "trailing" :317-326, get_variables :334-345,
get_mult :347-356, step :424-468, add_noise_to_xt :497-518.  Pinned by tests/golden/dpm_*.pt (outputs of the real class).

The per-frame scalar coefficients are the ones tokensgen_b200 uploads to the GPU (tg_cfg_dpm_step's `coef` table); the
tensor chains below reproduce torch's op-by-op rounding because they ARE torch ops on the same dtypes.
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np
import torch


class DpmTables:
    def __init__(self, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, snr_shift_scale=1.0,
                 rescale_betas_zero_snr=True):
        self.betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float64) ** 2
        alphas = 1.0 - self.betas
        ac = torch.cumprod(alphas, dim=0)
        ac = ac / (snr_shift_scale + (1 - snr_shift_scale) * ac)
        if rescale_betas_zero_snr:
            s = ac.sqrt()
            s0, sT = s[0].clone(), s[-1].clone()
            s = s - sT
            s = s * (s0 / (s0 - sT))
            ac = s ** 2
        self.alphas_cumprod = ac
        self.final_alpha_cumprod = torch.tensor(1.0)
        self.num_train_timesteps = num_train_timesteps

    def trailing_timesteps(self, num_inference_steps: int) -> np.ndarray:
        ratio = self.num_train_timesteps / num_inference_steps
        return np.round(np.arange(self.num_train_timesteps, 0, -ratio)).astype(np.int64) - 1

    def coefficients(self, timestep: int, prev_timestep: int, timestep_back: Optional[int]):
        """Scalars of one step as 0-dim fp64 tensors: (sqrt_alpha, sqrt_beta, mult0, mult1, mult2, mult3, mult_noise)."""
        a_t = self.alphas_cumprod[timestep]
        a_prev = self.alphas_cumprod[prev_timestep] if prev_timestep >= 0 else self.final_alpha_cumprod.double()
        a_back = self.alphas_cumprod[timestep_back] if timestep_back is not None else None
        lamb = ((a_t / (1 - a_t)) ** 0.5).log()
        lamb_next = ((a_prev / (1 - a_prev)) ** 0.5).log()
        h = lamb_next - lamb
        mult0 = ((1 - a_prev) / (1 - a_t)) ** 0.5 * (-h).exp()
        mult1 = (-2 * h).expm1() * a_prev ** 0.5
        if a_back is not None:
            lamb_prev = ((a_back / (1 - a_back)) ** 0.5).log()
            r = (lamb - lamb_prev) / h
            mult2, mult3 = 1 + 1 / (2 * r), 1 / (2 * r)
        else:
            mult2 = mult3 = torch.tensor(0.0, dtype=torch.float64)
        mult_noise = (1 - a_prev) ** 0.5 * (1 - (-2 * h).exp()) ** 0.5
        return a_t ** 0.5, (1 - a_t) ** 0.5, mult0, mult1, mult2, mult3, mult_noise


def _smul(c, x, device_semantics: str):
    """`c * x` for a 0-dim fp64 scalar tensor c and a tensor x, the way the reference writes it (scalar first).
    PyTorch evaluates this differently per device when x is bf16: the CUDA kernels keep the scalar in fp32 (opmath),
    the CPU kernels round it to bf16 first (only `x * c` keeps fp32 on CPU).  The reference runs on CUDA, so
    device_semantics="cuda" is the behaviour to reproduce; "cpu" is what the unmodified reference yields when it is run
    on the CPU to produce tests/golden/dpm.pt."""
    return c * x if device_semantics == "cpu" else x * c


def step(tables: DpmTables, model_output, old_x0, timestep, prev_timestep, timestep_back, sample, noise1, noise2,
         device_semantics: str = "cuda"):
    """CogVideoXDPMScheduler.step (v_prediction), scheduling_dpm_cogvideox.py:424-468, with the two randn draws
    supplied (noise1 = first draw, noise2 = second draw)."""
    sa, sb, m0, m1, m2, m3, mn = tables.coefficients(int(timestep), int(prev_timestep),
                                                     None if timestep_back is None else int(timestep_back))
    mul = lambda c, x: _smul(c, x, device_semantics)
    x0 = mul(sa, sample) - mul(sb, model_output)
    prev = mul(m0, sample) - mul(m1, x0) + mul(mn, noise1)
    if old_x0 is None or prev_timestep < 0:
        return prev, x0
    d = mul(m2, x0) - mul(m3, old_x0)
    return mul(m0, sample) - mul(m1, d) + mul(mn, noise2), x0


def window_step_bf16(tables: DpmTables, noise_pred, guidance_scale, latents, old_x0: List, t, prev_t, next_t, noise1, noise2,
                     device_semantics: str = "cuda", guidance_scale_img=None):
    """The FIFO worker's CFG + per-frame scheduler loop, cogvideo_sampling_mp_fifo.py:527-550 (all tensors bf16).
    noise_pred [2,F,...] — or [3,F,...] (uncond_txt, uncond_img, txt_img) with `guidance_scale_img`, the
    use_separate_guidance branch (:528-530) — latents [1,F,...], old_x0 list of F (tensor [1,1,...] or None);
    t/prev_t/next_t int arrays [F]."""
    if noise_pred.shape[0] == 3:
        ut, ui, a = noise_pred.chunk(3)
        npred = a + (guidance_scale - 1) * (a - ut) + (guidance_scale_img - 1) * (a - ui)
    else:
        u, c = noise_pred.chunk(2)
        npred = u + guidance_scale * (c - u)
    out = latents.clone()
    x0s = []
    for j in range(latents.shape[1]):
        back = int(next_t[j]) if next_t[j] > 0 else None
        p, x0 = step(tables, npred[:, [j]], old_x0[j], int(t[j]), int(prev_t[j]), back, latents[:, [j]],
                     noise1[:, [j]], noise2[:, [j]], device_semantics)
        out[:, [j]] = p.to(latents.dtype)
        x0s.append(x0.to(latents.dtype))
    return out, x0s


def add_noise_to_xt(tables: DpmTables, xt_prev, noise, timestep: int = 999):
    """scheduling_dpm_cogvideox.py:497-518: sqrt(1-beta_t) x + sqrt(beta_t) eps; the [1]-shaped fp64 scalars promote the
    expression to fp64, the caller's in-place assignment rounds it back to the latent dtype."""
    b = tables.betas[torch.tensor([timestep])]
    s1, s2 = (1 - b) ** 0.5, b ** 0.5
    while s1.dim() < xt_prev.dim():
        s1, s2 = s1.unsqueeze(-1), s2.unsqueeze(-1)
    return s1 * xt_prev + s2 * noise
