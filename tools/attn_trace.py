"""Developer tool: per-warp phase timeline of CTA (0,0) of the v2 attention kernel (clock64 stamps).
events per block: 0 loop top, 1 S ready, 2 S in registers, 3 row max exchanged, 4 exps+pack done, 5 P stored/arrived."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tokensgen_b200 import _ext as E  # noqa: E402

lib = E.load()
H, N = 8, 17776
nb = (N + 127) // 128
emu = int(sys.argv[1]) if len(sys.argv) > 1 else 0
stagger = int(sys.argv[2]) if len(sys.argv) > 2 else 0
E.set_tuning("attn_impl", 2); E.set_tuning("attn_emu", emu); E.set_tuning("attn_stagger", stagger)
q = torch.randn(1, H, N, 64, device="cuda").bfloat16()
k = torch.randn(1, H, N, 64, device="cuda").bfloat16()
v = torch.randn(1, H, N, 64, device="cuda").bfloat16()
out = torch.empty(1, N, H * 64, device="cuda", dtype=torch.bfloat16)
E.attn_fwd(q, k, v, out)
buf = torch.zeros(16, nb, 6, dtype=torch.int64, device="cuda")
lib.tg_debug_attn_trace(C.c_void_p(buf.data_ptr()))
E.attn_fwd(q, k, v, out)
torch.cuda.synchronize()
lib.tg_debug_attn_trace(C.c_void_p(0))
t = buf.cpu()
t0 = t[:, 0, 0].min()
t = t - t0
names = ["wait_S", "ld_S", "max+xchg", "exp+pack", "store_P"]
print(f"emu={emu} stagger={stagger}: total cycles for {nb} blocks: {int(t[:, -1, 5].max())}  per block {int(t[:, -1, 5].max()) / nb:.0f}")
for w in (0, 4, 8, 12):
    d = (t[w, 20:120, 1:] - t[w, 20:120, :-1]).float().mean(0)
    per = (t[w, 120, 0] - t[w, 20, 0]).item() / 100
    print(f"warp {w:2d} (tile {w // 8} half {(w // 4) % 2}): period {per:7.1f}  " + "  ".join(f"{n} {x:6.0f}" for n, x in zip(names, d.tolist())))
for j in range(60, 64):
    print(f"block {j}: " + " | ".join(f"w{w}: " + " ".join(str(int(x)) for x in (t[w, j] - t[0, 60, 0]).tolist()) for w in (0, 4, 8, 12)))
