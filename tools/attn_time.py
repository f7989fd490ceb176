"""Developer tool: full-size self-attention timing (2 x 48 heads x 17776^2 x 64) for a list of impl:emu:alt settings.
usage: [TG_LIB_PATH=...] python tools/attn_time.py 3:0:1 3:2:1 ..."""
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
for spec in sys.argv[1:]:
    impl, emu, alt = (int(x) for x in spec.split(":"))
    E.set_tuning("attn_impl", impl); E.set_tuning("attn_emu", emu); E.set_tuning("attn_alt", alt)
    for _ in range(2):
        E.attn_fwd(q, k, v, out)
    s, e = torch.cuda.Event(True), torch.cuda.Event(True)
    s.record()
    for _ in range(5):
        E.attn_fwd(q, k, v, out)
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 5
    err = ((out.float() - ref).norm() / ref.norm()).item()
    print(f"{os.path.basename(os.environ.get('TG_LIB_PATH', 'default'))} impl={impl} emu={emu} alt={alt}: {ms:.3f} ms {fl / ms / 1e9:.1f} TFLOP/s  rel_l2 vs sdpa {err:.2e}", flush=True)
