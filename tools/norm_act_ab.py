"""Developer tool: tg_vae_norm_act, shared-memory-staged kernel (dense output) vs register-fed kernel (pitched output) on the
decoder's activation shapes; 100 back-to-back launches each, CUDA events.  The staged kernel exists in the developer build only:
usage: TG_LIB_PATH=tokensgen_b200/libtokensgen_b200_dev.so TG_NORM_STAGED=1 python tools/norm_act_ab.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tokensgen_b200 import _ext as E  # noqa: E402

g = torch.Generator(device="cuda").manual_seed(0)
for (T, H, W, C, ratio) in ((8, 480, 720, 128, 8), (8, 240, 360, 256, 4), (4, 120, 180, 512, 2), (2, 60, 90, 512, 1), (8, 480, 720, 128, 0)):
    x = torch.randn(T, H, W, C, generator=g, device="cuda").bfloat16()
    gamma, beta = torch.ones(C, device="cuda").bfloat16(), torch.zeros(C, device="cuda").bfloat16()
    sums = E.vae_group_stats(x, 32)
    zy = zb = None
    if ratio:
        table = torch.randn(max(T // 4, 1), H // ratio, W // ratio, 2 * C, generator=g, device="cuda").bfloat16()
        zy, zb = table[..., :C], table[..., C:]
    dense = torch.empty(T, H, W, C, device="cuda", dtype=torch.bfloat16)
    pitched = torch.empty(T, H, W, C + 8, device="cuda", dtype=torch.bfloat16)[..., :C]
    res = {}
    for name, out in (("staged", dense), ("register-fed", pitched), ("staged", dense), ("register-fed", pitched)):
        for _ in range(20):
            E.vae_norm_act(x, sums, 32, 1e-6, gamma, beta, out, zy, zb, silu=True)
        s, e = torch.cuda.Event(True), torch.cuda.Event(True)
        s.record()
        for _ in range(100):
            E.vae_norm_act(x, sums, 32, 1e-6, gamma, beta, out, zy, zb, silu=True)
        e.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(e) / 100
        res.setdefault(name, []).append(ms)
    gb = x.numel() * 4 / 1e9
    print(f"[{T}x{H}x{W}x{C}] {'spatial' if ratio else 'plain'} {gb * 1e3:.0f} MB: " +
          "; ".join(f"{k} {min(v):.3f} ms = {gb / min(v) * 1e3:.0f} GB/s" for k, v in res.items()))
