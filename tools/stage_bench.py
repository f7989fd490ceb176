"""Measured times of the stages around the FIFO loop at their real sizes (random-init weights of the true shapes), one B200:
  * T2To DiT step: CogVideoX-5b-arch DiT with patch_size 1 on [2, 96, 16, 8, 12] (9 216 tokens + 226 text, plain processor)
  * Resampler: 13 x 1350 patch tokens of dim 3072 -> 4 x 8 x 12 condensed tokens (depth 4, 16 heads)
Prints JSON lines; combined with bench.py (window step) and tools/vae_bench.py (encode / decode) these give the per-stage
budget of configs[1..3] in profiles/r01_stage_times.md."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tokensgen_b200 import _ext as E  # noqa: E402
from tokensgen_b200.resampler import Resampler  # noqa: E402
from tokensgen_b200.rope import get_3d_rotary_pos_embed_v2  # noqa: E402
from tokensgen_b200.synth import build_random_model  # noqa: E402

dev = torch.device("cuda")
E.load()


def timed(fn, n=3, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(True), torch.cuda.Event(True)
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n


lin = lambda a, b, n: np.linspace(a, b, n, endpoint=False, dtype=np.float32)
with torch.no_grad():
    # ---- T2To step
    m = build_random_model(device=dev, seed=0, use_vip=False, patch_size=1)
    g = torch.Generator(device=dev).manual_seed(1)
    lat = torch.randn(2, 96, 16, 8, 12, generator=g, device=dev, dtype=torch.bfloat16)
    text = torch.randn(2, 226, 4096, generator=g, device=dev, dtype=torch.bfloat16)
    rope = get_3d_rotary_pos_embed_v2(64, lin(0, 96, 96), lin(0, 8, 8), lin(0, 12, 12), dim_t=52, dim_h=6, dim_w=6, device=dev)
    ts = torch.full((2,), 500, device=dev, dtype=torch.int64)
    ms = timed(lambda: m(lat, text, ts, image_rotary_emb=rope, return_dict=False))
    n = 9216 + 226
    flops = 2 * 42 * (2 * n * 3072 * (4 * 3072 + 2 * 12288) + 4 * 48 * n * n * 64)
    print(json.dumps({"stage": "T2To DiT step (B=2, 9442 tokens, 42 layers, no vip)", "ms": round(ms, 1), "tflop": round(flops / 1e12, 1),
                      "tflops": round(flops / ms / 1e9, 1), "steps_per_video": 52}), flush=True)
    del m
    torch.cuda.empty_cache()
    # ---- Resampler
    r = Resampler(dim=3072, depth=4, dim_head=64, heads=16, num_height_queries=8, num_width_queries=12, num_temporal_queries=4,
                  embedding_dim=3072, output_dim=3072, max_height_seq_len=30, max_width_seq_len=45, max_temporal_seq_len=13)
    for p in r.parameters():
        torch.nn.init.normal_(p, std=0.02)
    r = r.to(dev, torch.bfloat16).eval()
    x = torch.randn(1, 13, 1350, 3072, generator=g, device=dev, dtype=torch.bfloat16)
    img = get_3d_rotary_pos_embed_v2(64, lin(0, 13, 13), lin(0, 30, 30), lin(0, 45, 45), device=dev)
    smp = get_3d_rotary_pos_embed_v2(64, lin(1000, 1013, 4), lin(0, 30, 8), lin(0, 45, 12), device=dev)
    ms = timed(lambda: r(x, image_rotary_emb=img, sampling_rotary_emb=smp))
    print(json.dumps({"stage": "Resampler, one 13-frame chunk (17 550 tokens -> 384 condensed tokens)", "ms": round(ms, 2),
                      "calls_per_video": "chunks + 1"}), flush=True)
