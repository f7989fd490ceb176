"""BASELINE.json configs[2]/[3] in miniature: the FIFO stage of gen.yaml (CogVideoX-5b-shaped random-init DiT with the
video-IP-adapter, 52-step diagonal queue, 13x30x45 windows, CFG pair) run for real through `cogvideo_fifo_mp_v2` on P ranks
with NCCL boundary exchange, for `chunks` 13-frame chunks instead of 24.  The base-stage bundle is synthesised (random
latents / embeddings of the true shapes) so that the GPU time goes into the FIFO stage itself.
Prints one JSON line: wall time, per-iteration times, denoised latent tokens/s over the whole stage and in steady state.
usage: torchrun --nproc-per-node P tools/fifo_full_run.py [chunks=2]"""
import json
import os
import sys
import time
from types import SimpleNamespace

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tokensgen_b200 import _ext as E  # noqa: E402
from tokensgen_b200.fifo import FifoSchedule, cogvideo_fifo_mp_v2  # noqa: E402
from tokensgen_b200.pipeline import FIFOCogVideoXPipelineOutput  # noqa: E402
from tokensgen_b200.rope import get_3d_rotary_pos_embed, vip_position_grids  # noqa: E402
from tokensgen_b200.scheduler import CogVideoXDPMScheduler  # noqa: E402
from tokensgen_b200.synth import build_random_model  # noqa: E402

chunks = int(sys.argv[1]) if len(sys.argv) > 1 else 2
world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
E.load()
nf, T = 13, 52
model = build_random_model(device=dev, seed=0)          # same weights on every rank (as after loading one checkpoint)
sch = CogVideoXDPMScheduler.cogvideox_5b()
sch.set_timesteps(T)
g = torch.Generator(device=dev).manual_seed(1)
mk = lambda *s: torch.randn(*s, generator=g, device=dev, dtype=torch.bfloat16)
img_grid, cond_grid = vip_position_grids(60, 90, 2, chunks, nf, 4, 8, 12, 1000)
base = FIFOCogVideoXPipelineOutput(
    fifo_latents=mk(1, T, 16, 60, 90), fifo_old_pred_original_sample=[mk(1, 1, 16, 60, 90) for _ in range(T - 1)] + [None],
    orig_latents=mk(1, nf, 16, 60, 90), nf_per_chunk=nf, vip_nf_per_chunk=4, num_frames=chunks * nf,
    image_embeddings=mk(2, 4 * (chunks + 1), 3072, 8, 12), timesteps=sch.timesteps, num_inference_steps=T,
    do_classifier_free_guidance=True, use_separate_guidance=False, use_dynamic_cfg=False, prompt_embeds=mk(2, 226, 4096),
    image_rotary_emb=get_3d_rotary_pos_embed(64, [[0, 0, 0], [nf, 30, 45]], (nf, 30, 45), device=dev),
    vip_image_rotary_grid=list(img_grid), vip_condition_rotary_grid=list(cond_grid), cache_idx=[], guidance_scale=6.0,
    guidance_scale_img=6.0, extra_step_kwargs={}, video_ipadapter_start_frame_idx=1000, sampling_params={"num_partitions": 4},
    output_type="latent", return_dict=False)
pipe = SimpleNamespace(transformer=model, scheduler=sch)
sched = FifoSchedule(chunks * nf, [int(t) for t in sch.timesteps], nf, 4, True)
stamps = []


def tick(it):
    torch.cuda.synchronize()
    stamps.append(time.perf_counter())


if world > 1:
    dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
with torch.no_grad():
    _, latents, _ = cogvideo_fifo_mp_v2([pipe], base, seed=7, progress=tick)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
wall = time.perf_counter() - t0
if rank == 0:
    its = np.diff(np.array([t0] + stamps))
    n_win = [len(sched.windows(i)) for i in range(sched.num_iterations)]
    full = [dt for dt, n in zip(its, n_win) if n == 8]
    line = {"what": "FIFO stage (gen.yaml shapes), real run", "n_gpus": world, "chunks": chunks, "iterations": sched.num_iterations,
            "window_forwards": int(sum(n_win)), "wall_s": round(wall, 2), "emitted_latent_frames": int(latents.shape[1]),
            "denoised_latent_tokens_per_s": round(1350 * latents.shape[1] / wall, 1),
            "steady_state_iteration_s": round(float(np.median(full)), 4) if full else None,
            "steady_state_tokens_per_s": round(1350 / float(np.median(full)), 1) if full else None,
            "first_iteration_s": round(float(its[0]), 3), "latents_finite": bool(torch.isfinite(latents.float()).all()),
            "rounds_per_iteration_at_8_windows": -(-8 // world)}
    print(json.dumps(line), flush=True)
if world > 1:
    dist.destroy_process_group()
