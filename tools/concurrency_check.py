"""Developer / release check: full-size DiT window forwards on the main stream WHILE the tiled VAE coder (tiles on
TG_VAE_TILE_STREAMS streams) runs on a side stream — the combination the streaming decode produces under the FIFO loop
(CTA-pair GEMMs, the attention launches + K6 side stream, CTA-pair / tap-reuse convolutions, normalise passes, all concurrent).
Prints a line per finished round; run under `timeout -k 5 ...` (a device hang does not answer SIGINT).
usage: python tools/concurrency_check.py [rounds]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from vae_bench import build_vae  # noqa: E402
from tokensgen_b200 import vae as V  # noqa: E402
from tokensgen_b200.rope import get_3d_rotary_pos_embed, get_3d_rotary_pos_embed_v2  # noqa: E402
from tokensgen_b200.synth import build_random_model, window_inputs  # noqa: E402

rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 3
dev = torch.device("cuda", 0)
model = build_random_model(device=dev, seed=0)
vae = build_vae()
vae.enable_tiling()
host = window_inputs(seed=42)
inp = {k: v.to(dev) for k, v in host.items()}
F = 13
rope = get_3d_rotary_pos_embed(64, [[0, 0, 0], [F, 30, 45]], (F, 30, 45), device=dev)
img_rope = get_3d_rotary_pos_embed_v2(64, np.arange(F, dtype=np.float32) + 45, np.arange(30, dtype=np.float32),
                                      np.arange(45, dtype=np.float32), device=dev)
cond_rope = get_3d_rotary_pos_embed_v2(64, np.array([1000, 1003.25, 1006.5, 1009.75, 1013], dtype=np.float32),
                                       np.linspace(0, 30, 8, endpoint=False, dtype=np.float32),
                                       np.linspace(0, 45, 12, endpoint=False, dtype=np.float32), device=dev)
ts = torch.full((2, F), 500, device=dev)
g = torch.Generator().manual_seed(1)
z = torch.randn(1, 16, 13, 60, 90, generator=g).cuda().bfloat16()
x = (torch.rand(1, 3, 49, 480, 720, generator=g) * 2 - 1).cuda().bfloat16()
side = torch.cuda.Stream()
print(f"tile streams = {V._TILE_STREAMS}", flush=True)


def dit():
    lat = inp["latents"]
    return model(hidden_states=torch.cat([lat, lat]), encoder_hidden_states=inp["prompt_embeds"], timestep=ts,
                 vip_encoder_hidden_states=inp["image_embeddings"], image_rotary_emb=rope, vip_image_rotary_emb=img_rope,
                 vip_condition_rotary_emb=cond_rope, return_dict=False)[0]


with torch.no_grad():
    ref_dit, ref_dec, ref_enc = dit(), vae.decode(z).sample, vae.encode(x).latent_dist.parameters   # each alone
    torch.cuda.synchronize()
    for r in range(rounds):
        t0 = time.time()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            decs = [vae.decode(z).sample for _ in range(2)]
            encs = [vae.encode(x).latent_dist.parameters for _ in range(2)]
        outs = [dit() for _ in range(2)]
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        same = (all(torch.equal(o, ref_dit) for o in outs) and all(torch.equal(d, ref_dec) for d in decs)
                and all(torch.equal(e, ref_enc) for e in encs))
        print(f"round {r}: ok {time.time() - t0:.2f} s, identical to the stand-alone results: {same}", flush=True)
