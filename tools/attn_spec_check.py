"""Developer tool: speculative softmax reference (attn_spec 1) vs the exact running-max path (attn_spec 0) on the full-size
self-attention (2 x 48 heads x 17776^2 x 64): accuracy vs SDPA, 5-launch burst and sustained (power-capped) timings.
usage: python tools/attn_spec_check.py [--sustained]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tokensgen_b200 import _ext as E  # noqa: E402

B, H, N = 2, 48, 17776
torch.manual_seed(0)
q = torch.randn(B, H, N, 64, device="cuda").bfloat16()
k = torch.randn(B, H, N, 64, device="cuda").bfloat16()
v = torch.randn(B, H, N, 64, device="cuda").bfloat16()
out = torch.empty(B, N, H * 64, device="cuda", dtype=torch.bfloat16)
ref = torch.nn.functional.scaled_dot_product_attention(q, k, v).permute(0, 2, 1, 3).flatten(2).float()
fl = 4 * B * H * N * N * 64


def timed(fn, n):
    s, e = torch.cuda.Event(True), torch.cuda.Event(True)
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n


sdpa = lambda: torch.nn.functional.scaled_dot_product_attention(q, k, v)
for _ in range(2):
    sdpa()
print(f"torch SDPA burst: {timed(sdpa, 5):.3f} ms", flush=True)
for spec in (1, 0, 1, 0):
    for emu in (0, 1):
        E.set_tuning("attn_spec", spec); E.set_tuning("attn_emu", emu)
        for _ in range(2):
            E.attn_fwd(q, k, v, out)
        ms = timed(lambda: E.attn_fwd(q, k, v, out), 5)
        err = ((out.float() - ref).norm() / ref.norm()).item()
        print(f"spec={spec} emu={emu} burst: {ms:.3f} ms {fl / ms / 1e9:.1f} TFLOP/s rel_l2 vs sdpa {err:.2e}", flush=True)
if "--sustained" in sys.argv:
    E.set_tuning("attn_emu", 0)
    for name, fn in (("sdpa", sdpa), ("spec=1", None), ("spec=0", None), ("spec=1", None)):
        if fn is None:
            E.set_tuning("attn_spec", int(name[-1]))
            fn = lambda: E.attn_fwd(q, k, v, out)
        for _ in range(120):
            fn()
        ms = timed(fn, 120)
        print(f"{name} sustained: {ms:.3f} ms {fl / ms / 1e9:.1f} TFLOP/s", flush=True)
