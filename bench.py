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


# ------------------------------------------------------------------------------------------------------ CPU baseline
def cpu_block_seconds(repeats: int, warmup: int):
    """The reference's algorithm on the host cores: ONE CogVideoX-5b block + VIP at full size (B = 1, 17 550 video + 226
    text + 480 vip tokens) through the oracle port (PyTorch CPU eager, bf16 like the reference), all cores."""
    from oracle import dit as odit
    from oracle import rope as orope
    from oracle.synth import dit_shapes, synth_state_dict
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    shapes = {k: v for k, v in dit_shapes(48, 64, 1, 512, 4096, 16, 16, 2, 3072, True).items()
              if k.startswith("transformer_blocks.0.")}
    sd = synth_state_dict(shapes, 77)
    g = torch.Generator().manual_seed(42)
    hid = torch.randn(1, TOKENS, 3072, generator=g).bfloat16()
    enc = torch.randn(1, 706, 3072, generator=g).bfloat16()
    temb = torch.randn(1, 13, 512, generator=g).bfloat16()
    rope = orope.window_rope(64, 13, 30, 45)
    img = orope.rope_3d_from_grids(64, np.arange(13, dtype=np.float32), np.arange(30, dtype=np.float32), np.arange(45, dtype=np.float32))
    cond = orope.rope_3d_from_grids(64, np.array([1000, 1003.25, 1006.5, 1009.75, 1013], dtype=np.float32),
                                    np.linspace(0, 30, 8, endpoint=False, dtype=np.float32),
                                    np.linspace(0, 45, 12, endpoint=False, dtype=np.float32))
    cfg = odit.DitConfig()
    times = []
    with torch.no_grad():
        for i in range(warmup + repeats):
            t0 = time.perf_counter()
            odit.block_forward(sd, "transformer_blocks.0", cfg, hid, enc, temb, rope, img, cond, torch.bfloat16)
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
    return times, cores


def eager_block_seconds(device, repeats: int = 5, warmup: int = 2):
    """Second baseline: the reference's op sequence in PyTorch eager ON THE GPU — the oracle port
    executes the same torch ops the reference modules do (18 cuBLAS Linears, three SDPA calls, unfused LayerNorm / modulation
    / RoPE / cat / gated residual) — for ONE CogVideoX-5b block + VIP at full size (B = 1), bf16.  Not the product path."""
    from oracle import dit as odit
    from oracle import rope as orope
    from oracle.synth import dit_shapes, synth_state_dict
    dev = torch.device(device)
    shapes = {k: v for k, v in dit_shapes(48, 64, 1, 512, 4096, 16, 16, 2, 3072, True).items()
              if k.startswith("transformer_blocks.0.")}
    sd = {k: v.to(dev, torch.bfloat16) for k, v in synth_state_dict(shapes, 77).items()}
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
    cfg = odit.DitConfig()
    sync = torch.cuda.synchronize if dev.type == "cuda" else (lambda: None)
    times = []
    with torch.no_grad():
        for i in range(warmup + repeats):
            sync()
            t0 = time.perf_counter()
            odit.block_forward(sd, "transformer_blocks.0", cfg, hid, enc, temb, rope, img, cond, torch.bfloat16)
            sync()
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    return float(np.median(times))


def cpu_tokens_per_s(block_seconds: float) -> float:
    # one bench step = 42 blocks x CFG pair (B = 2) of that block (embedding/final layers are < 0.1 % and left out)
    return TOKENS / (DENOISE_STEPS * 42 * 2 * block_seconds)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    times, cores = cpu_block_seconds(args.steps, args.warmup)
    sec = float(np.mean(times))
    val = cpu_tokens_per_s(sec)
    sample = (f"{args.steps} timed + {args.warmup} warm-up forwards of ONE CogVideoX-5b block + VIP at full size (B=1) through the "
              f"oracle port (PyTorch CPU eager bf16); a step = 42 blocks x 2 CFG branches, extrapolated")
    line = {"impl": "reference", "metric": "denoised latent tokens/sec", "value": val, "unit": "tokens/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 84 * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "denoise_steps_per_clip": DENOISE_STEPS},
            "cpu_baseline": {"value": val, "unit": "tokens/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


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

    if rank == 0:
        hbm, tf_burst, tf_sust, src = measured_peaks()
        per = {k: sum(s.elapsed_time(e) for s, e in v) / args.steps for k, v in prof.items()}  # ms per step per op tag
        n_calls = {k: len(v) // args.steps for k, v in prof.items()}
        top = max(per, key=per.get)
        self_attn = f"attn_fwd[q{TOKENS + 226},kv{TOKENS + 226}]"
        dom = self_attn if self_attn in per else top
        N = TOKENS + 226
        flops_per_launch = 4 * 2 * 48 * N * N * 64
        avg_ms = per[dom] / n_calls[dom]
        achieved = flops_per_launch / avg_ms / 1e9
        ms_step = ms_total / args.steps
        value = world * TOKENS * args.steps / (DENOISE_STEPS * ms_total / 1e3)
        e2e = world * TOKENS * args.steps / (DENOISE_STEPS * ms_e2e / 1e3)
        line = {"metric": "denoised latent tokens/sec", "value": value, "unit": "tokens/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "bf16", "data": "synthetic",
                "config": {"workload": WORKLOAD, "denoise_steps_per_clip": DENOISE_STEPS, "token_steps_per_s": value * DENOISE_STEPS,
                           "model_tflops_per_step": 775.9, "achieved_model_tflops": 775.9 / (ms_step / 1e3) ,
                           "l2": "inputs larger than L2: 14.3 GB of weights + 3 GB of activations stream through every step",
                           "parallelism": f"window-parallel x{world}" + (" + NCCL boundary-frame exchange" if world > 1 else "")},
                "e2e": {"value": e2e, "unit": "tokens/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                        "ms_per_step": ms_e2e / args.steps},
                "gpu_launches": launches,
                "roofline": {"bound": "tensor", "kernel": "attn_fwd_kernel (self-attention, 2x48 heads x 17776^2 x 64)",
                             "achieved": achieved, "peak": tf_sust * 1.0, "unit": "TFLOP/s", "frac": achieved / tf_sust,
                             "peak_source": f"{src} bf16_tflops_sustained (kernel timed inside a long step)",
                             "frac_of_burst_peak": achieved / tf_burst, "avg_launch_ms": avg_ms,
                             "share_of_step": per[dom] / ms_step, "traffic": attn_dram_traffic(),
                             "traffic_unit": "bytes per launch (ncu dram read+write, profiles/r01_attn3_full_summary.md)",
                             "algorithmic_bytes": 4 * 2 * 48 * N * 64 * 2},
                "kernel_ms_per_step": {k: round(v, 3) for k, v in sorted(per.items(), key=lambda kv: -kv[1])},
                "clocks": clocks}
        if seqpar is not None:
            seqpar["speedup_vs_this_runs_1gpu_step"] = ms_step / seqpar["ms_per_step"]
            line["sequence_parallel"] = seqpar
        if world == 1 and not args.no_eager_baseline:
            # reported next to the CPU baseline, never on the product path; a failure here must not cost the bench line
            try:
                sec = eager_block_seconds(dev)
                line["gpu_eager_baseline"] = {
                    "what": "the reference's op sequence in PyTorch eager on this GPU (oracle port: cuBLAS Linears + SDPA + unfused "
                            "elementwise), ONE CogVideoX-5b block + VIP at full size, B = 1, bf16; a step = 42 blocks x 2 CFG branches",
                    "block_ms": sec * 1e3, "value": cpu_tokens_per_s(sec), "unit": "tokens/s",
                    "ours_over_eager": value / cpu_tokens_per_s(sec)}
            except Exception as e:  # noqa: BLE001
                line["gpu_eager_baseline"] = {"error": f"{type(e).__name__}: {e}"[:300]}
        if world == 1 and not args.no_cpu_baseline:
            times, cores = cpu_block_seconds(1, 1)
            line["cpu_baseline"] = {"value": cpu_tokens_per_s(times[0]), "unit": "tokens/s", "cores": cores, "kind": "port",
                                    "sample": "1 timed + 1 warm-up forward of ONE CogVideoX-5b block + VIP at full size (B=1) through "
                                              "the oracle port (PyTorch CPU eager bf16), extrapolated x42 blocks x2 CFG branches",
                                    "block_seconds": times[0]}
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
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        if args.warmup < 3:
            args.warmup = 3
        run_ours(args)


if __name__ == "__main__":
    main()
