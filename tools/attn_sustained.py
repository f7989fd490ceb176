"""Developer tool: sustained (power-capped) self-attention timing — the full-size call looped for ~3 s, second half timed.
usage: [TG_LIB_PATH=...] python tools/attn_sustained.py [emu]"""
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
ref = torch.nn.functional.scaled_dot_product_attention(q[:, :4], k[:, :4], v[:, :4]).permute(0, 2, 1, 3).flatten(2).float()
E.set_tuning("attn_emu", int(sys.argv[1]) if len(sys.argv) > 1 else 0)
for _ in range(150):
    E.attn_fwd(q, k, v, out)
s, e = torch.cuda.Event(True), torch.cuda.Event(True)
s.record()
for _ in range(150):
    E.attn_fwd(q, k, v, out)
e.record()
torch.cuda.synchronize()
ms = s.elapsed_time(e) / 150
err = ((out[:, :, :256].float() - ref).norm() / ref.norm()).item()
print(f"{os.path.basename(os.environ.get('TG_LIB_PATH', 'default'))} sustained: {ms:.3f} ms {4 * B * H * N * N * 64 / ms / 1e9:.1f} TFLOP/s rel_l2 {err:.2e}", flush=True)
