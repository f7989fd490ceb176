"""Developer tool: run every stage of the tiny pipeline twice on identical inputs and report which ones are not bit-stable."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_pipeline_gpu import _tiny_pipe  # noqa: E402

pipe = _tiny_pipe()
g = torch.Generator().manual_seed(3)
frames = (torch.rand(1, 3, 9, 64, 96, generator=g) * 2 - 1).cuda().bfloat16()


def same(name, fn, n=4):
    outs = [fn() for _ in range(n)]
    torch.cuda.synchronize()
    bad = [i for i in range(1, n) if not torch.equal(outs[0], outs[i])]
    md = max(((outs[0].float() - o.float()).abs().max().item() for o in outs[1:]), default=0.0)
    print(f"{name:28s} {'stable' if not bad else 'UNSTABLE'}  max|d|={md:.3e}")
    return outs[0]


with torch.no_grad():
    par = same("vae.encode", lambda: pipe.vae.encode(frames).latent_dist.parameters.clone())
    z = par[:, :16].contiguous()
    same("vae.decode", lambda: pipe.vae.decode(z).sample.clone())
    pipe.vae.enable_tiling()
    same("vae.decode tiled", lambda: pipe.vae.decode(z).sample.clone())
    same("vae.encode tiled", lambda: pipe.vae.encode(frames).latent_dist.parameters.clone())
    pipe.vae.disable_tiling()
    tok = torch.randn(1, 3, 24, 256, generator=g).cuda().bfloat16()
    from oracle import rope as orope
    import numpy as np
    lin = lambda a, b, n: np.linspace(a, b, n, endpoint=False, dtype=np.float32)
    ir = orope.rope_3d_from_grids(64, lin(0, 3, 3), lin(0, 4, 4), lin(0, 6, 6))
    sr = orope.rope_3d_from_grids(64, lin(1000, 1003, 2), lin(0, 4, 2), lin(0, 6, 3))
    emb = same("resampler", lambda: pipe.resampler(tok, image_rotary_emb=ir, sampling_rotary_emb=sr).clone())
    lat = torch.randn(2, 3, 16, 8, 12, generator=g).cuda().bfloat16()
    text = torch.randn(2, 10, 128, generator=g).cuda().bfloat16()
    vip = torch.randn(2, 3, 256, 2, 3, generator=g).cuda().bfloat16()
    ts = torch.tensor([[999, 640, 21], [999, 640, 21]]).cuda()
    rope = orope.rope_3d(64, [[0, 0, 0], [3, 4, 6]], (3, 4, 6))
    cr = orope.rope_3d_from_grids(64, lin(1000, 1004.5, 3), lin(0, 4, 2), lin(0, 6, 3))
    same("dit forward", lambda: pipe.transformer(lat, text, ts, vip_encoder_hidden_states=vip, image_rotary_emb=rope,
                                                 vip_image_rotary_emb=ir, vip_condition_rotary_emb=cr, return_dict=False)[0].clone())
