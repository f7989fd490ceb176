"""Sequence-parallel (Ulysses) forward over real peers: run under torchrun with P >= 2 ranks.

  --tiny : tiny DiT (4 heads, 2 layers, with the video-IP-adapter): the P-rank forward must be bit-identical to the
           single-GPU forward on every rank; prints SEQPAR_OK.
  --full : CogVideoX-5b shapes at the bench geometry (CFG pair, 17 550 + 226 + 480 rows): bit-identity of the sharded forward,
           then the latency of the full 42-layer forward, sharded vs unsharded (CUDA events, max over ranks), one JSON line.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tiny", action="store_true")
    ap.add_argument("--full", action="store_true")
    ap.add_argument("--layers", type=int, default=42)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    import tokensgen_b200.transformer as T
    from tokensgen_b200 import _ext as E
    from tokensgen_b200.rope import get_3d_rotary_pos_embed, get_3d_rotary_pos_embed_v2
    from tokensgen_b200.synth import build_random_model, window_inputs
    T._FUSE_PAIR = True  # the sharded path fuses K4 + K5; the unsharded comparison uses the same kernel sequence

    def tmax(ms):
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    def timed(fn, n):
        dist.barrier(); torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(n):
            fn()
        e.record()
        dist.barrier(); torch.cuda.synchronize()
        return tmax(s.elapsed_time(e) / n)

    ok = True
    if args.tiny:
        F, Hh, Ww = 3, 8, 12
        vip_kw = dict(length=12, func_type="1", scale=[0.6],
                      resampler_params=dict(output_dim=128, num_height_queries=2, num_width_queries=3, num_temporal_queries=1))
        m = build_random_model(device=dev, seed=5, vip_kwargs=vip_kw, num_attention_heads=4, time_embed_dim=128,
                               text_embed_dim=128, num_layers=2)
        g = torch.Generator().manual_seed(9)
        lat = torch.randn(2, F, 16, Hh, Ww, generator=g).bfloat16().to(dev)
        text = torch.randn(2, 10, 128, generator=g).bfloat16().to(dev)
        vip = torch.randn(2, 2, 128, 2, 3, generator=g).bfloat16().to(dev)
        ts = torch.tensor([[900., 800., 700.]] * 2, device=dev)
        rope = get_3d_rotary_pos_embed(64, [[0, 0, 0], [F, Hh // 2, Ww // 2]], (F, Hh // 2, Ww // 2), device=dev)
        img = get_3d_rotary_pos_embed_v2(64, np.arange(F, dtype=np.float32), np.arange(Hh // 2, dtype=np.float32),
                                         np.arange(Ww // 2, dtype=np.float32), device=dev)
        cond = get_3d_rotary_pos_embed_v2(64, np.array([1000., 1003.], dtype=np.float32), np.array([0., 2.], dtype=np.float32),
                                          np.array([0., 2., 4.], dtype=np.float32), device=dev)
        call = lambda: m(lat, text, ts, vip_encoder_hidden_states=vip, image_rotary_emb=rope, vip_image_rotary_emb=img,
                         vip_condition_rotary_emb=cond, return_dict=False)[0]
        with torch.no_grad():
            ref = call().clone()
            m.enable_sequence_parallel()
            for _ in range(3):
                out = call()
            torch.cuda.synchronize()
        same = torch.equal(out, ref)
        print(f"[{rank}] tiny: sharded over {world} ranks bit-identical = {same}", flush=True)
        ok = ok and same

    if args.full:
        m = build_random_model(device=dev, seed=0, num_layers=args.layers)
        host = window_inputs(seed=42)
        F = 13
        rope = get_3d_rotary_pos_embed(64, [[0, 0, 0], [F, 30, 45]], (F, 30, 45), device=dev)
        img = get_3d_rotary_pos_embed_v2(64, np.arange(F, dtype=np.float32) + 45, np.arange(30, dtype=np.float32),
                                         np.arange(45, dtype=np.float32), device=dev)
        cond = get_3d_rotary_pos_embed_v2(64, np.array([1000, 1003.25, 1006.5, 1009.75, 1013], dtype=np.float32),
                                          np.linspace(0, 30, 8, endpoint=False, dtype=np.float32),
                                          np.linspace(0, 45, 12, endpoint=False, dtype=np.float32), device=dev)
        lat = host["latents"].to(dev)
        lat2 = torch.cat([lat, lat])
        prompt, vip = host["prompt_embeds"].to(dev), host["image_embeddings"].to(dev)
        ts = torch.linspace(900, 500, F, device=dev).expand(2, -1).contiguous()
        call = lambda: m(lat2, prompt, ts, vip_encoder_hidden_states=vip, image_rotary_emb=rope, vip_image_rotary_emb=img,
                         vip_condition_rotary_emb=cond, return_dict=False)[0]
        with torch.no_grad():
            ref = call().clone()
            for _ in range(2):
                call()
            ms_one = timed(call, args.steps)
            E.profile = {}
            call(); torch.cuda.synchronize()
            prof_one, E.profile = E.profile, None
            m.enable_sequence_parallel()
            out = call().clone()
            same = torch.equal(out, ref)
            for _ in range(2):
                call()
            l0 = E.launch_count
            ms_sp = timed(call, args.steps)
            launches = (E.launch_count - l0) // args.steps
            E.profile = {}
            call(); torch.cuda.synchronize()
            prof_sp, E.profile = E.profile, None
        per = lambda prof: {k: round(sum(s_.elapsed_time(e_) for s_, e_ in v), 3) for k, v in
                            sorted(prof.items(), key=lambda kv: -sum(s_.elapsed_time(e_) for s_, e_ in kv[1]))[:10]}
        ok = ok and same
        if rank == 0:
            line = {"what": "CogVideoX-5b DiT forward, CFG pair, 18 256 rows, sequence-parallel (Ulysses, fused peer-store all-to-alls)",
                    "layers": args.layers, "ranks": world, "bit_identical_to_single_gpu": same,
                    "ms_single_gpu": ms_one, "ms_sharded": ms_sp, "speedup": ms_one / ms_sp,
                    "parallel_efficiency": ms_one / ms_sp / world, "launches_per_forward": launches,
                    "op_ms_single_gpu": per(prof_one), "op_ms_sharded_rank0": per(prof_sp)}
            print(json.dumps(line), flush=True)
            if args.out:
                with open(args.out, "a") as f:
                    f.write(json.dumps(line) + "\n")
    dist.barrier()
    if rank == 0:
        print("SEQPAR_OK" if ok else "SEQPAR_MISMATCH", flush=True)
    ok_t = torch.tensor([int(ok)], device=dev)
    dist.all_reduce(ok_t, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    sys.exit(0 if ok_t.item() == 1 else 1)


if __name__ == "__main__":
    main()
