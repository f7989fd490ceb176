"""Small driver for ncu captures of the attention kernel (the bench launch: CFG pair x 48 heads x 17 776 tokens).
usage: attn_profile.py [heads] [impl] [emu] [stagger]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tokensgen_b200 import _ext as E  # noqa: E402

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
