"""Small driver for ncu captures of the attention kernel (the bench launch: CFG pair x 48 heads x 17 776 tokens).
usage: attn_profile.py [heads] [impl] [emu] [stagger]      (knobs: developer build only)
       attn_profile.py pair                                 the SHIPPED launch: self-attention + vip cross-attention in one
                                                            launch (tg_attn_fwd_pair)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tokensgen_b200 import _ext as E  # noqa: E402

if len(sys.argv) > 1 and sys.argv[1] == "pair":
    B, H, N, n_vip = 2, 48, 17776, 480
    torch.manual_seed(0)
    mk = lambda n: torch.randn(B, H, n, 64, device="cuda").bfloat16()
    q, k, v = mk(N), mk(N), mk(N)
    q2, k2, v2 = mk(N + n_vip), mk(N + n_vip), mk(N + n_vip)
    out = torch.empty(B, N + n_vip, H * 64, device="cuda", dtype=torch.bfloat16)
    for _ in range(3):
        E.attn_fwd_pair(q, k, v, N, N, q2, k2, v2, N, n_vip, out, 0.6015625)
    torch.cuda.synchronize()
    sys.exit(0)
H = int(sys.argv[1]) if len(sys.argv) > 1 else 16
if len(sys.argv) > 2:
    E.set_tuning("attn_impl", int(sys.argv[2]))
if len(sys.argv) > 3:
    E.set_tuning("attn_emu", int(sys.argv[3]))
if len(sys.argv) > 4:
    E.set_tuning("attn_stagger", int(sys.argv[4]))
N = 17776
torch.manual_seed(0)
B = 2
q = torch.randn(B, H, N, 64, device="cuda").bfloat16()
k = torch.randn(B, H, N, 64, device="cuda").bfloat16()
v = torch.randn(B, H, N, 64, device="cuda").bfloat16()
out = torch.empty(B, N, H * 64, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    E.attn_fwd(q, k, v, out)
torch.cuda.synchronize()
