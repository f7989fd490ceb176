"""One-off parity check at BASELINE's full sizes: the CogVideoX-5b-shaped DiT (48 heads x 64, 226 text + 17 550 video + 480 vip
tokens, CFG pair, per-frame timesteps) with L layers, CUDA mirror vs the fp32 oracle on the same bf16-rounded weights and
inputs — shows how the bf16 error grows with depth.  Too slow for pytest (the fp32 oracle needs ~10 s per layer on the box's
CPU cores).  usage: python tools/full_size_parity.py [layers ...]   (default 1 2 4)"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import dit as odit  # noqa: E402
from oracle import rope as orope  # noqa: E402
from oracle.synth import dit_shapes, synth_state_dict  # noqa: E402
from tokensgen_b200.transformer import CogVideoXTransformer3DModel  # noqa: E402

layers = [int(a) for a in sys.argv[1:]] or [1, 2, 4]
torch.set_num_threads(os.cpu_count() or 1)
g = torch.Generator().manual_seed(42)
lat = torch.randn(1, 13, 16, 60, 90, generator=g).bfloat16()
lat2 = torch.cat([lat, lat])
text = torch.randn(2, 226, 4096, generator=g).bfloat16()
vip = torch.randn(2, 5, 3072, 8, 12, generator=g).bfloat16()
ts = torch.tensor([[999 - 19 * i for i in range(13)]] * 2)
rope = orope.window_rope(64, 13, 30, 45)
img = orope.rope_3d_from_grids(64, np.arange(13, dtype=np.float32) + 45, np.arange(30, dtype=np.float32), np.arange(45, dtype=np.float32))
cond = orope.rope_3d_from_grids(64, np.array([1000, 1003.25, 1006.5, 1009.75, 1013], dtype=np.float32),
                                np.linspace(0, 30, 8, endpoint=False, dtype=np.float32), np.linspace(0, 45, 12, endpoint=False, dtype=np.float32))
for L in layers:
    sd = synth_state_dict(dit_shapes(48, 64, L, 512, 4096, 16, 16, 2, 3072, True), 100 + L)
    m = CogVideoXTransformer3DModel(num_attention_heads=48, attention_head_dim=64, time_embed_dim=512, text_embed_dim=4096,
                                    num_layers=L, use_rotary_positional_embeddings=True, attention_bias=True)
    m.set_vip_layers(None, length=480, func_type="1", scale=[0.6],
                     resampler_params=dict(output_dim=3072, num_height_queries=8, num_width_queries=12, num_temporal_queries=4))
    m.load_state_dict(sd, strict=True)
    m = m.to("cuda", torch.bfloat16).eval()
    with torch.no_grad():
        y = m(lat2.cuda(), text.cuda(), ts.cuda(), vip_encoder_hidden_states=vip.cuda(), image_rotary_emb=rope,
              vip_image_rotary_emb=img, vip_condition_rotary_emb=cond, return_dict=False)[0].float().cpu()
    t0 = time.time()
    cfg = odit.DitConfig(num_layers=L)
    with torch.no_grad():
        ref = odit.dit_forward(sd, cfg, lat2, text, ts, vip, rope, img, cond, torch.float32)
    err = ((y - ref).norm() / ref.norm()).item()
    print(json.dumps({"layers": L, "rel_l2_vs_fp32_oracle": err, "max_abs": (y - ref).abs().max().item(),
                      "ref_rms": ref.pow(2).mean().sqrt().item(), "oracle_seconds": round(time.time() - t0, 1)}), flush=True)
    del m
    torch.cuda.empty_cache()
