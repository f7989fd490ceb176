#!/usr/bin/env python
"""Benchmark of the TokensGen FIFO-denoising hot path on B200 (contract: see the task statement / DESIGN.md §Measurement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one denoise step of one 13x30x45 latent window (17 550 video tokens): the CogVideoX-5b DiT forward for the
classifier-free-guidance pair (B = 2, per-frame timesteps, 226 text + 480 condensed VIP tokens, 42 layers) followed by
the fused CFG + per-frame DPM-Solver++ step — i.e. fifo_onestep_per_gpu of the reference, which is also one step of the
To2V base stage (BASELINE.json configs[1]).  metric = denoised latent tokens/s = 17 550 * steps / (52 * seconds): a clip's
tokens are denoised after the 52-step schedule the shipped configs use (config/infer/edit.yaml:8).

  value : device-resident inputs, CUDA-event timed, max over ranks.
  e2e   : the same step through the public API with HOST (pinned) window inputs copied in and results copied out
          every step — what the reference's controller<->worker queue traffic is.
  N > 1 : one process per GPU (torchrun), each rank denoises its own window (weak scaling) and exchanges the FIFO
          boundary frames with its ring neighbours through NCCL every step.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

TOKENS = 13 * 30 * 45
DENOISE_STEPS = 52
WORKLOAD = ("configs[1] To2V edit.yaml single clip: 13x30x45 latent window (49 frames 480x720), CFG pair B=2, "
            "CogVideoX-5b DiT 42 layers + video-IP-adapter (480 condensed tokens), one denoise step per bench step")


def attn_dram_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE self-attention launch of this shape, from the committed
    `ncu --set full` capture summary (profiles/r01_attn3_full_summary.md); None if the summary is missing."""
    import re
    path = os.path.join(ROOT, "profiles", "r02_attn_pair_full_summary.md")     # the shipped pair launch (round 2)
    if not os.path.exists(path):
        path = os.path.join(ROOT, "profiles", "r01_attn3_full_summary.md")
    if not os.path.exists(path):
        return None
    txt = open(path).read()
    tot = 0.0
    for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        m = re.search(r"\| " + re.escape(key) + r" \| ([0-9.]+) \| (\w+) \|", txt)
        if not m:
            return None
        tot += float(m.group(1)) * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[m.group(2)]
    heads = re.search(r"self-attention, (\d+)x48 heads", txt)
    scale = 2.0 / float(heads.group(1)) if heads else 1.0   # the capture may hold one CFG branch; a bench launch has two
    return tot * scale


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), "measured"
    return 6650.0, 1590.0, 1400.0, "fallback"


class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.FIELDS}",
                                       "--format=csv,noheader,nounits", "-lms", "200"], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        rows = [l.strip().split(", ") for l in open(self.f.name) if l.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[3:7]):
                if v.strip().lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------ reference block
def _block_inputs(dev):
    """one CogVideoX-5b block + VIP at full size, B = 1, 17 550 video + 226 text + 480 vip tokens."""
    from oracle import rope as orope
    g = torch.Generator().manual_seed(42)
    hid = torch.randn(1, TOKENS, 3072, generator=g).bfloat16().to(dev)
    enc = torch.randn(1, 706, 3072, generator=g).bfloat16().to(dev)
    temb = torch.randn(1, 13, 512, generator=g).bfloat16().to(dev)
    on = lambda pair: tuple(t.to(dev) for t in pair)
    rope = on(orope.window_rope(64, 13, 30, 45))
    img = on(orope.rope_3d_from_grids(64, np.arange(13, dtype=np.float32), np.arange(30, dtype=np.float32),
                                      np.arange(45, dtype=np.float32)))
    cond = on(orope.rope_3d_from_grids(64, np.array([1000, 1003.25, 1006.5, 1009.75, 1013], dtype=np.float32),
                                       np.linspace(0, 30, 8, endpoint=False, dtype=np.float32),
                                       np.linspace(0, 45, 12, endpoint=False, dtype=np.float32)))
    return hid, enc, temb, rope, img, cond


def reference_block_runner(device):
    """A callable running ONE forward of the reference's CogVideoXBlock + video-IP-adapter (func_type "1") at full size in bf16
    on `device`, and which implementation it is:
      "reference" — the reference's OWN module (cogvideox_transformer_3d.py:54-332 + attention_processor.py:1955-2155),
                    unmodified, from baseline/_ref (built by oracle/vendor_reference.py; imports through the stubs);
      "port"      — the oracle restatement (oracle/dit.py::block_forward), only when baseline/_ref is absent.
    Both get the same seeded weights (oracle.synth.synth_state_dict(shapes, 77)) and inputs."""
    from oracle import vendor_reference as vr
    from oracle.synth import dit_shapes, synth_state_dict
    dev = torch.device(device)
    shapes = {k: v for k, v in dit_shapes(48, 64, 1, 512, 4096, 16, 16, 2, 3072, True).items()
              if k.startswith("transformer_blocks.0.")}
    sd = synth_state_dict(shapes, 77)
    hid, enc, temb, rope, img, cond = _block_inputs(dev)
    if vr.enable():
        from longvgen.models.cogvideox_transformer_3d import CogVideoXBlock
        blk = CogVideoXBlock(dim=3072, num_attention_heads=48, attention_head_dim=64, time_embed_dim=512, attention_bias=True)
        blk.set_vip_layers(length=480, func_type="1", scale=[0.6])
        missing = blk.load_state_dict({k[len("transformer_blocks.0."):]: v for k, v in sd.items()}, strict=True)
        blk = blk.to(dev, torch.bfloat16).eval()

        def run():
            return blk(hid, enc, temb, image_rotary_emb=rope, vip_image_rotary_emb=img, vip_condition_rotary_emb=cond)
        return run, "reference"
    from oracle import dit as odit
    sd = {k: v.to(dev, torch.bfloat16) for k, v in sd.items()}
    cfg = odit.DitConfig()

    def run_port():
        return odit.block_forward(sd, "transformer_blocks.0", cfg, hid, enc, temb, rope, img, cond, torch.bfloat16)
    return run_port, "port"


def cpu_block_seconds(repeats: int, warmup: int):
    """The reference's CPU implementation of the path on the host cores: ONE block forward per sample, all cores."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    run, kind = reference_block_runner("cpu")
    times = []
    with torch.no_grad():
        for i in range(warmup + repeats):
            t0 = time.perf_counter()
            run()
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
    return times, cores, kind


def eager_block_seconds(device, repeats: int = 5, warmup: int = 2):
    """Second baseline: the same reference block in PyTorch eager ON THE GPU (cuBLAS Linears, three
    SDPA-flash calls, unfused LayerNorm / modulation / RoPE / cat / gated residual), bf16.  Not the product path."""
    run, kind = reference_block_runner(device)
    times = []
    with torch.no_grad():
        for i in range(warmup + repeats):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            run()
            torch.cuda.synchronize()
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    return float(np.median(times)), kind


def cpu_tokens_per_s(block_seconds: float) -> float:
    # one bench step = 42 blocks x CFG pair (B = 2) of that block (embedding/final layers are < 0.1 % and left out)
    return TOKENS / (DENOISE_STEPS * 42 * 2 * block_seconds)


def make_config(world: int, value: float, ms_step: float):
    """`config` of the JSON line — the same keys on both arms (the driver compares them)."""
    return {"workload": WORKLOAD, "denoise_steps_per_clip": DENOISE_STEPS, "token_steps_per_s": value * DENOISE_STEPS,
            "model_tflops_per_step": 775.9, "achieved_model_tflops": 775.9 / (ms_step / 1e3),
            "l2": "inputs larger than L2: 14.3 GB of weights + 3 GB of activations stream through every step",
            "parallelism": f"window-parallel x{world}" + (" + NCCL boundary-frame exchange" if world > 1 else "")}


def run_reference(args):
    """`--impl reference`: the reference's own CPU implementation of the path on this box's host cores.  One STEP of this arm
    is a bounded SAMPLE of a bench step — one of its 42 x 2 block forwards (a whole step is ~2.5 min of CPU): `ms_per_step`
    is the measured time of that sample, `value` the metric it implies for the whole step (x 84), `sample_fraction` says so."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    times, cores, kind = cpu_block_seconds(args.steps, args.warmup)
    sec = float(np.mean(times))
    val = cpu_tokens_per_s(sec)
    what = ("the reference's own CogVideoXBlock + VideoIPAdapterCogVideoXAttnProcessor2_0 (unmodified, baseline/_ref)" if kind == "reference"
            else "the oracle port of the reference block (baseline/_ref absent)")
    sample = (f"{args.steps} timed + {args.warmup} warm-up forwards of ONE CogVideoX-5b block + VIP at full size (B=1), {what}, PyTorch "
              f"CPU eager bf16, {cores} threads; a bench step = 42 blocks x 2 CFG branches = 84 such samples")
    line = {"impl": "reference", "metric": "denoised latent tokens/sec", "value": val, "unit": "tokens/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": make_config(1, val, sec * 84 * 1e3),
            "sample_fraction": 1.0 / 84, "ms_per_whole_step_extrapolated": sec * 84 * 1e3,
            "cpu_baseline": {"value": val, "unit": "tokens/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": val, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)



# ------------------------------------------------------------------------------------------------------ real FIFO stage
def fifo_schedule_stats(chunks: int, world: int):
    """Pure index arithmetic of the FIFO stage of a `chunks`-chunk video (gen.yaml: 24): iterations, window forwards, the
    rounds a P-rank job needs (max windows on one rank, per iteration) and the speed-up bound that follows."""
    from tokensgen_b200.fifo import FifoSchedule
    from tokensgen_b200.scheduler import CogVideoXDPMScheduler
    sch = CogVideoXDPMScheduler.cogvideox_5b()
    sch.set_timesteps(DENOISE_STEPS)
    s = FifoSchedule(chunks * 13, [int(t) for t in sch.timesteps], 13, 4, True)
    wins = [s.windows(i) for i in range(s.num_iterations)]
    forwards = sum(len(w) for w in wins)
    rounds = sum(max(sum(1 for x in w if x.rank % world == r) for r in range(world)) for w in wins)
    ramp = sum(1 for w in wins if len(w) < 8)
    # with RampSharding: an iteration whose `a` active windows allow groups of g = 2^k ranks (a * g <= P) costs 1 / g of a step
    sizes = [g for g in (8, 4, 2) if g <= world and world % g == 0]
    sharded = 0.0
    for w in wins:
        g = max([g for g in sizes if len(w) * g <= world] or [1])
        sharded += 1.0 / g if g > 1 else max(sum(1 for x in w if x.rank % world == r) for r in range(world))
    return {"iterations": s.num_iterations, "window_forwards": forwards, "rounds": rounds, "ramp_iterations": ramp,
            "schedule_bound_speedup": forwards / rounds, "rounds_with_ramp_sharding": sharded,
            "bound_speedup_with_ramp_sharding": forwards / sharded}


def run_fifo_stage(model, sch, dev, world, rank, chunks: int, barrier):
    """BASELINE.json configs[2]/[3] for real: the FIFO stage of a `chunks`-chunk gen.yaml video (CogVideoX-5b shapes, 52-step
    diagonal queue, 13 x 30 x 45 windows, CFG pair, video-IP-adapter, lookahead write-back) through the product's own
    `cogvideo_fifo_mp_v2` controller — window rank w on process w % P, NCCL boundary-frame exchange every iteration.  The
    priming bundle is synthetic (random latents / embeddings of the true shapes).  Timed on the device, max over ranks."""
    import torch.distributed as dist
    from types import SimpleNamespace
    from tokensgen_b200.fifo import cogvideo_fifo_mp_v2
    from tokensgen_b200.pipeline import FIFOCogVideoXPipelineOutput
    from tokensgen_b200.rope import get_3d_rotary_pos_embed, vip_position_grids
    nf, T = 13, DENOISE_STEPS
    g = torch.Generator(device=dev).manual_seed(1)
    mk = lambda *s_: torch.randn(*s_, generator=g, device=dev, dtype=torch.bfloat16)
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

    class _WarmedUp(Exception):
        pass

    def stop_after_ramp_groups(it):   # iterations 0..7 touch every ramp-sharding group size (8, 4, 2 ranks per window)
        if it == 7:
            raise _WarmedUp()

    # untimed warm-up, like the W warm-up steps of the window bench: the first use of each sequence-parallel group allocates
    # and rendezvouses its peer-mapped workspaces (once per process, not per video)
    import copy
    try:
        with torch.no_grad():
            cogvideo_fifo_mp_v2([pipe], copy.copy(base), seed=7, progress=stop_after_ramp_groups)
    except _WarmedUp:
        pass
    barrier()
    s_ev, e_ev = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s_ev.record()
    with torch.no_grad():
        _, latents, _ = cogvideo_fifo_mp_v2([pipe], base, seed=7)
    e_ev.record()
    barrier()
    ms = s_ev.elapsed_time(e_ev)
    if world > 1:
        tt = torch.tensor([ms], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = tt.item()
    return ms / 1e3, int(latents.shape[1]), bool(torch.isfinite(latents.float()).all())

def vae_block(sweep_frames):
    """The `vae` object of the N = 1 line (runs in the child process started by run_vae_block)."""
    from tools import vae_bench as vb
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    vae = vb.build_vae()
    pts = [vb.point(vae, "decode", 13, False, peaks), vb.point(vae, "decode", 13, True, peaks),
           vb.point(vae, "encode", 49, False, peaks), vb.point(vae, "encode", 49, True, peaks)]
    sweep = [vb.point(vae, "decode", T, False, peaks) for T in sweep_frames]
    return {"what": "3D causal VAE, full CogVideoX-5b widths (128,256,256,512), 480x720: decode of one 13-latent-frame "
                    "clip (49 frames) untiled and tiled 3x3 (the reference CLI's default), encode of 49 frames; "
                    "`sweep`: configs[4] decode of T latent frames as ONE causal stream",
            "points": pts, "sweep": [{k: p[k] for k in ("latent_frames", "pixel_frames", "ms", "pixel_frames_per_s",
                                                         "tflops_whole_pass", "frac_of_sustained_peak")} for p in sweep]}


def run_vae_block(sweep_frames, timeout_s: float):
    """vae_block() in a child process (its own CUDA context on the same GPU), killed with its process group at the time limit."""
    import signal
    import subprocess
    cmd = [sys.executable, os.path.abspath(__file__), "--vae-child", "--vae-sweep"] + [str(t) for t in sweep_frames]
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT")}
    try:
        p = subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, cwd=ROOT, env=env, start_new_session=True)
    except OSError as e:
        return {"error": f"{type(e).__name__}: {e}"[:300]}
    try:
        out, err = p.communicate(timeout=timeout_s)
    except subprocess.TimeoutExpired:
        try:
            os.killpg(p.pid, signal.SIGKILL)     # a process blocked on a hung device does not answer SIGINT / SIGTERM
        except ProcessLookupError:
            pass
        p.wait()
        return {"error": f"the VAE block did not finish within {timeout_s:.0f} s and was killed"}
    lines = [l for l in out.splitlines() if l.startswith("{")]
    if p.returncode != 0 or not lines:
        return {"error": f"child exited with {p.returncode}: {err.strip()[-300:]}"}
    try:
        return json.loads(lines[-1])
    except ValueError as e:
        return {"error": f"{type(e).__name__}: {e}"[:300]}


# ------------------------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import torch.distributed as dist
    from tokensgen_b200 import _ext as E
    from tokensgen_b200.rope import get_3d_rotary_pos_embed, get_3d_rotary_pos_embed_v2
    from tokensgen_b200.scheduler import CogVideoXDPMScheduler
    from tokensgen_b200.synth import build_random_model, window_inputs

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (ours): needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    E.load()
    for key in ("attn_spec", "attn_emu"):      # A/B knobs of the DEVELOPER build only (TG_LIB_PATH=.../libtokensgen_b200_dev.so)
        if os.environ.get("TG_" + key.upper()) is not None:
            E.set_tuning(key, int(os.environ["TG_" + key.upper()]))

    model = build_random_model(device=dev, seed=rank)
    sch = CogVideoXDPMScheduler.cogvideox_5b()
    sch.set_timesteps(DENOISE_STEPS)
    host = window_inputs(seed=42 + rank)
    F = 13
    rope = get_3d_rotary_pos_embed(64, [[0, 0, 0], [F, 30, 45]], (F, 30, 45), device=dev)
    img_rope = get_3d_rotary_pos_embed_v2(64, np.arange(F, dtype=np.float32) + 45, np.arange(30, dtype=np.float32),
                                          np.arange(45, dtype=np.float32), device=dev)
    cond_rope = get_3d_rotary_pos_embed_v2(64, np.array([1000, 1003.25, 1006.5, 1009.75, 1013], dtype=np.float32),
                                           np.linspace(0, 30, 8, endpoint=False, dtype=np.float32),
                                           np.linspace(0, 45, 12, endpoint=False, dtype=np.float32), device=dev)
    # a steady-state FIFO window: 13 consecutive entries of the 52-step schedule, all with x0 history
    ts_all = sch.timesteps.numpy()
    t = ts_all[20:33][::-1].copy()
    prev_t = ts_all[21:34][::-1].copy()
    next_t = ts_all[19:32][::-1].copy()
    ts_dev = torch.as_tensor(t, device=dev).expand(2, -1).contiguous()

    dev_in = {k: v.to(dev) for k, v in host.items()}
    prompt = dev_in["prompt_embeds"]
    out_host = {"latents": torch.empty_like(host["latents"]).pin_memory(), "x0": torch.empty_like(host["old_x0"]).pin_memory()}
    right, left = (rank + 1) % world, (rank - 1) % world
    frame = dev_in["latents"][0, 0]
    send_l = torch.empty((2, 7) + tuple(frame.shape), device=dev, dtype=torch.bfloat16)
    recv_l = torch.empty_like(send_l)
    send_r = torch.empty((2, 1) + tuple(frame.shape), device=dev, dtype=torch.bfloat16)
    recv_r = torch.empty_like(send_r)

    def exchange(lat, x0):
        """FIFO boundary exchange (tokensgen_b200.fifo.run_fifo's transfers at P = world): 7 frames to the right
        neighbour, 1 to the left, latents + x0 history, grouped NCCL send/recv."""
        send_l[0].copy_(lat[0, 6:13]); send_l[1].copy_(x0[6:13])
        send_r[0].copy_(lat[0, 6:7]); send_r[1].copy_(x0[6:7])
        ops = [dist.P2POp(dist.isend, send_l, right), dist.P2POp(dist.irecv, recv_l, left),
               dist.P2POp(dist.isend, send_r, left), dist.P2POp(dist.irecv, recv_r, right)]
        for r in dist.batch_isend_irecv(ops):
            r.wait()

    def step(inputs):
        lat = inputs["latents"]
        noise_pred = model(hidden_states=torch.cat([lat, lat]), encoder_hidden_states=prompt, timestep=ts_dev,
                           vip_encoder_hidden_states=inputs["image_embeddings"], image_rotary_emb=rope,
                           vip_image_rotary_emb=img_rope, vip_condition_rotary_emb=cond_rope, return_dict=False)[0]
        old = [inputs["old_x0"][j].unsqueeze(0).unsqueeze(0) for j in range(F)]
        out_lat, x0s = sch.window_step(noise_pred, lat, old, t, prev_t, next_t, 6.0, noise=(inputs["noise1"], inputs["noise2"]))
        x0 = torch.cat([x.reshape((1,) + tuple(frame.shape)) for x in x0s])
        if world > 1:
            exchange(out_lat, x0)
        return out_lat, x0

    h2d_keys = ("latents", "old_x0", "image_embeddings")
    h2d_bytes = sum(host[k].numel() * 2 for k in h2d_keys)
    d2h_bytes = sum(v.numel() * 2 for v in out_host.values())

    def step_e2e():
        inputs = dict(dev_in)
        for k in h2d_keys:
            inputs[k] = host[k].to(dev, non_blocking=True)
        out_lat, x0 = step(inputs)
        out_host["latents"].copy_(out_lat, non_blocking=True)
        out_host["x0"].copy_(x0, non_blocking=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(n):
            fn()
        e.record()
        barrier()
        ms = s.elapsed_time(e)
        if world > 1:
            tt = torch.tensor([ms], device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = tt.item()
        return ms

    with torch.no_grad():
        for _ in range(args.warmup):
            step(dev_in)
        sampler = ClockSampler(local) if rank == 0 else None
        E.profile = {}
        launches0 = E.launch_count
        ms_total = timed(lambda: step(dev_in), args.steps)
        launches = E.launch_count - launches0
        prof, E.profile = E.profile, None
        for _ in range(min(args.warmup, 2)):
            step_e2e()
        ms_e2e = timed(step_e2e, args.steps)
        clocks = sampler.stop() if sampler else None
        # N > 1, extra (not part of `value`): the SAME window step with the DiT forward sequence-parallel over all N ranks
        # (tokensgen_b200/seqpar.py: rows sharded for LayerNorm / GEMMs, heads for attention, all-to-alls fused into the
        # epilogues as NVLink peer stores) — the strong-scaling latency of ONE clip, which is what the serial base stage
        # of the pipeline uses.  Every rank runs rank 0's window (same weights, same inputs).
        seqpar = None
        if world > 1 and 48 % world == 0 and not args.no_seqpar:
            del model
            torch.cuda.empty_cache()
            model = build_random_model(device=dev, seed=0)
            sp_in = {k: v.to(dev) for k, v in window_inputs(seed=42).items()}
            sp_prompt = sp_in["prompt_embeds"]

            def step_sp():
                lat = sp_in["latents"]
                noise_pred = model(hidden_states=torch.cat([lat, lat]), encoder_hidden_states=sp_prompt, timestep=ts_dev,
                                   vip_encoder_hidden_states=sp_in["image_embeddings"], image_rotary_emb=rope,
                                   vip_image_rotary_emb=img_rope, vip_condition_rotary_emb=cond_rope, return_dict=False)[0]
                old = [sp_in["old_x0"][j].unsqueeze(0).unsqueeze(0) for j in range(F)]
                return sch.window_step(noise_pred, lat, old, t, prev_t, next_t, 6.0, noise=(sp_in["noise1"], sp_in["noise2"]))[0]

            import tokensgen_b200.transformer as T
            T._FUSE_PAIR = True   # same kernel sequence (K4 + K5 in one launch) on both sides of the bit-identity check
            ref_out = step_sp().clone()
            model.enable_sequence_parallel()
            for _ in range(args.warmup):
                out_sp = step_sp()
            same = bool(torch.equal(out_sp, ref_out))
            ms_sp = timed(step_sp, args.steps) / args.steps
            model.disable_sequence_parallel()
            T._FUSE_PAIR = False
            seqpar = {"what": "one window step (DiT forward CFG pair + DPM step) sharded over all ranks: strong scaling of a single clip",
                      "ranks": world, "ms_per_step": ms_sp, "bit_identical_to_unsharded": same}

        # N > 1: the REAL FIFO stage (configs[2]/[3]) through the controller — ramp-up (fewer active windows than GPUs for the
        # first 40 iterations), steady state, boundary exchange, emit / shift / re-noise — on a short video
        fifo = None
        if world > 1 and not args.no_fifo_stage:
            chunks = args.fifo_chunks
            est = lambda c: fifo_schedule_stats(c, world)["rounds"] * (ms_total / args.steps) / 1e3
            while chunks > 1 and est(chunks) > args.fifo_budget_s:
                chunks -= 1
            wall, emitted, finite = run_fifo_stage(model, sch, dev, world, rank, chunks, barrier)
            st = fifo_schedule_stats(chunks, world)
            one_gpu = st["window_forwards"] * (ms_total / args.steps) / 1e3
            gen = fifo_schedule_stats(24, world)
            fifo = {"what": f"FIFO stage of a {chunks}-chunk gen.yaml video ({chunks * 13} latent frames = {chunks * 49} video frames of "
                            "480x720) through cogvideo_fifo_mp_v2: real controller, NCCL boundary exchange, device-timed, max over ranks",
                    "chunks": chunks, **st, "wall_s": wall, "emitted_latent_frames": emitted, "latents_finite": finite,
                    # the schedule costs `rounds_with_ramp_sharding` window-step equivalents on P ranks (a ramp iteration whose
                    # windows run on groups of g ranks counts 1 / g); the measured wall over that count is the cost of one such
                    # step inside the stage (window step + boundary exchange + emit / shift), and the bound is forwards / count
                    "emitted_tokens_per_s": 1350 * emitted / wall, "s_per_step_equivalent": wall / st["rounds_with_ramp_sharding"],
                    "one_gpu_counterpart_s": one_gpu,
                    "one_gpu_counterpart_is": "window_forwards x this run's measured single-window step time (one GPU runs the windows "
                                              "of an iteration back to back); profiles/ holds a measured 1-GPU run of the same stage",
                    "speedup_vs_one_gpu": one_gpu / wall,
                    "efficiency_vs_schedule_bound": (one_gpu / wall) / st["bound_speedup_with_ramp_sharding"],
                    "gen_yaml_24_chunks": {**gen, "projected_wall_s": gen["rounds_with_ramp_sharding"] * wall / st["rounds_with_ramp_sharding"],
                                           "projected_speedup_vs_one_gpu": gen["bound_speedup_with_ramp_sharding"] * (one_gpu / wall)
                                           / st["bound_speedup_with_ramp_sharding"]}}

    if rank == 0:
        hbm, tf_burst, tf_sust, src = measured_peaks()
        per = {k: sum(s.elapsed_time(e) for s, e in v) / args.steps for k, v in prof.items()}  # ms per step per op tag
        n_calls = {k: len(v) // args.steps for k, v in prof.items()}
        N, n_vip = TOKENS + 226, 480
        # the dominant kernel: the self-attention launch — alone (tg_attn_fwd) or, by default, with the text/video -> vip
        # cross-attention folded in as a second pass over the same query tiles (tg_attn_fwd_pair)
        self_attn, pair = f"attn_fwd[q{N},kv{N}]", f"attn_fwd_pair[q{N},kv{N}+kv{n_vip}]"
        if pair in per:
            dom, flops_per_launch = pair, 4 * 2 * 48 * 64 * (N * N + N * n_vip)
            dom_name = f"attn3_fwd_kernel, pair launch (self-attention 2x48 heads x {N}^2 x 64 + cross-attention to {n_vip} vip keys)"
            alg_bytes = (4 * N + N + 2 * n_vip) * 2 * 48 * 64 * 2          # q, k, v, out + q2 (N rows) + k2, v2 (480 rows each), bf16
        else:
            dom = self_attn if self_attn in per else max(per, key=per.get)
            flops_per_launch = 4 * 2 * 48 * N * N * 64
            dom_name = f"attn3_fwd_kernel (self-attention, 2x48 heads x {N}^2 x 64)"
            alg_bytes = 4 * 2 * 48 * N * 64 * 2
        avg_ms = per[dom] / n_calls[dom]
        achieved = flops_per_launch / avg_ms / 1e9
        ms_step = ms_total / args.steps
        value = world * TOKENS * args.steps / (DENOISE_STEPS * ms_total / 1e3)
        e2e = world * TOKENS * args.steps / (DENOISE_STEPS * ms_e2e / 1e3)
        line = {"metric": "denoised latent tokens/sec", "value": value, "unit": "tokens/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "bf16", "data": "synthetic",
                "config": make_config(world, value, ms_step),
                "e2e": {"value": e2e, "unit": "tokens/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                        "ms_per_step": ms_e2e / args.steps},
                "gpu_launches": launches,
                "roofline": {"bound": "tensor", "kernel": dom_name, "flops_per_launch": flops_per_launch,
                             "achieved": achieved, "peak": tf_sust * 1.0, "unit": "TFLOP/s", "frac": achieved / tf_sust,
                             "peak_source": f"{src} bf16_tflops_sustained (kernel timed inside a long step)",
                             "frac_of_burst_peak": achieved / tf_burst, "avg_launch_ms": avg_ms,
                             "share_of_step": per[dom] / ms_step, "traffic": attn_dram_traffic(),
                             "traffic_unit": "bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum of one launch; "
                                             "profiles/r02_attn_pair_full_summary.md, else round 1's capture of the self-attention alone)",
                             "algorithmic_bytes": alg_bytes},
                "kernel_ms_per_step": {k: round(v, 3) for k, v in sorted(per.items(), key=lambda kv: -kv[1])},
                "clocks": clocks}
        if seqpar is not None:
            seqpar["speedup_vs_this_runs_1gpu_step"] = ms_step / seqpar["ms_per_step"]
            line["sequence_parallel"] = seqpar
        if fifo is not None:
            line["fifo_stage"] = fifo
        line["value_definition"] = ("17 550 tokens x steps / (52 x seconds): every window step advances 13 latent frames by one of the 52 "
                                    "denoise levels (the To2V base stage, configs[1]); in the FIFO stage a window step FINALISES only its "
                                    "second half (lookahead), so emitted tokens/s there is fifo_stage.emitted_tokens_per_s (N > 1 lines)")
        if world == 1 and not args.no_eager_baseline:
            # reported next to the CPU baseline, never on the product path; a failure here must not cost the bench line
            try:
                sec, kind = eager_block_seconds(dev)
                line["gpu_eager_baseline"] = {
                    "what": "the reference block in PyTorch eager on this GPU (cuBLAS Linears + SDPA-flash + unfused elementwise), ONE "
                            "CogVideoX-5b block + VIP at full size, B = 1, bf16; a step = 42 blocks x 2 CFG branches",
                    "kind": kind, "block_ms": sec * 1e3, "value": cpu_tokens_per_s(sec), "unit": "tokens/s",
                    "ours_over_eager": value / cpu_tokens_per_s(sec)}
            except Exception as e:  # noqa: BLE001
                line["gpu_eager_baseline"] = {"error": f"{type(e).__name__}: {e}"[:300]}
        if world == 1 and not args.no_vae:
            # configs[4] + the VAE bookends of configs[1]: full-size CogVideoX VAE, random-init weights, this GPU.  Guarded like the
            # eager baseline (a failure is recorded, not raised) and run in a CHILD process under a time limit: the block must
            # never cost the bench line, not even by not returning
            del model
            torch.cuda.empty_cache()
            line["vae"] = run_vae_block(args.vae_sweep, args.vae_timeout_s)
        if world == 1 and not args.no_cpu_baseline:
            times, cores, kind = cpu_block_seconds(3, 1)
            sec = float(np.median(times))
            line["cpu_baseline"] = {"value": cpu_tokens_per_s(sec), "unit": "tokens/s", "cores": cores, "kind": kind,
                                    "sample": "3 timed (median) + 1 warm-up forward of ONE CogVideoX-5b block + VIP at full size (B=1): "
                                              + ("the reference's own module from baseline/_ref" if kind == "reference" else "the oracle port")
                                              + ", PyTorch CPU eager bf16, all host threads; a step = 84 such forwards (42 blocks x 2 "
                                                "CFG branches), extrapolated",
                                    "block_seconds": sec}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-eager-baseline", action="store_true", help="skip the PyTorch-eager-on-GPU baseline block")
    ap.add_argument("--no-seqpar", action="store_true", help="N > 1: skip the extra sequence-parallel single-clip measurement")
    ap.add_argument("--no-fifo-stage", action="store_true", help="N > 1: skip the real FIFO-stage run")
    ap.add_argument("--fifo-chunks", type=int, default=3, help="N > 1: length of the FIFO-stage video in 13-frame chunks")
    ap.add_argument("--fifo-budget-s", type=float, default=320.0, help="N > 1: shorten the FIFO video until it fits this many seconds")
    ap.add_argument("--no-vae", action="store_true", help="N = 1: skip the VAE block")
    ap.add_argument("--vae-sweep", type=int, nargs="*", default=[25, 49], help="N = 1: extra decode points (latent frames, one stream)")
    ap.add_argument("--vae-timeout-s", type=float, default=300.0, help="N = 1: time limit of the VAE block (child process)")
    ap.add_argument("--vae-child", action="store_true", help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.vae_child:
        try:
            print(json.dumps(vae_block(args.vae_sweep)), flush=True)
        except Exception as e:  # noqa: BLE001
            print(json.dumps({"error": f"{type(e).__name__}: {e}"[:300]}), flush=True)
        return
    if args.impl == "reference":
        run_reference(args)
    else:
        if args.warmup < 3:
            args.warmup = 3
        run_ours(args)


if __name__ == "__main__":
    main()
