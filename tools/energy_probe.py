"""Developer tool: run one hot kernel back to back for a few seconds while sampling nvidia-smi, to get its equilibrium power,
SM clock and time per call on this box -> energy per call.  The window step is power-capped (sw_power_cap, ~1550 MHz), so
energy per step, not idle time, is what sets the step time.
usage: energy_probe.py   (prints one line per kernel)"""
import os
import subprocess
import sys
import tempfile
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tokensgen_b200 import _ext as E  # noqa: E402

dev = "cuda"
torch.manual_seed(0)
B, F, hw, n_text, n_vip, d, H = 2, 13, 1350, 226, 480, 3072, 48
n_video = F * hw
rows = n_text + n_video + n_vip
M = B * rows
rm = E.make_rowmap(n_text, n_video, n_vip, hw, F)
table = torch.randn(B * F, 18 * d, device=dev).bfloat16()


def sample(fn, seconds=4.0):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
    pr = subprocess.Popen(["nvidia-smi", "-i", "0", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-lms", "100"],
                          stdout=f, stderr=subprocess.DEVNULL)
    s, e = torch.cuda.Event(True), torch.cuda.Event(True)
    t0 = time.time()
    n = 0
    s.record()
    while time.time() - t0 < seconds:
        for _ in range(10):
            fn()
        n += 10
        torch.cuda.synchronize()
    e.record()
    torch.cuda.synchronize()
    pr.terminate()
    pr.wait()
    rows_ = [l.split(",") for l in open(f.name) if "," in l]
    os.unlink(f.name)
    half = rows_[len(rows_) // 2:]                      # second half: thermal / power equilibrium
    clk = float(np.median([float(r[0]) for r in half]))
    pw = float(np.median([float(r[1]) for r in half]))
    ms = s.elapsed_time(e) / n
    return ms, clk, pw


a = torch.randn(M, d, device=dev).bfloat16()
w1 = (torch.randn(4 * d, d, device=dev) / d ** 0.5).bfloat16()
b1 = torch.randn(4 * d, device=dev).bfloat16()
hbuf = torch.empty(M, 4 * d, device=dev, dtype=torch.bfloat16)
w2 = (torch.randn(d, 4 * d, device=dev) / (4 * d) ** 0.5).bfloat16()
b2 = torch.randn(d, device=dev).bfloat16()
x = torch.randn(M, d, device=dev).bfloat16()
gate = E.make_modvec(table[:, 5 * d:6 * d], table[:, 2 * d:3 * d], table[:, 14 * d:15 * d])
N = n_text + n_video
q = torch.randn(B, H, N, 64, device=dev).bfloat16()
k = torch.randn(B, H, N, 64, device=dev).bfloat16()
v = torch.randn(B, H, N, 64, device=dev).bfloat16()
o = torch.empty(B, N, H * 64, device=dev, dtype=torch.bfloat16)
lnw, lnb = torch.ones(d, device=dev).bfloat16(), torch.zeros(d, device=dev).bfloat16()
shift = E.make_modvec(table[:, 3 * d:4 * d], table[:, 0:d], table[:, 12 * d:13 * d])
scale = E.make_modvec(table[:, 4 * d:5 * d], table[:, d:2 * d], table[:, 13 * d:14 * d])
y = torch.empty_like(x)
cases = [("self-attention 2x48x17776^2 (7.77 TFLOP)", lambda: E.attn_fwd(q, k, v, o), 7.766),
         ("FF1 GEMM + GELU 36512x12288x3072 (2.76 TFLOP)", lambda: E.gemm_bias_act(a, w1, b1, hbuf, act=E.ACT_GELU_TANH), 2.757),
         ("FF2 GEMM + gate residual 36512x3072x12288 (2.76 TFLOP)", lambda: E.gemm_gate_residual(hbuf, w2, b2, x, B, rm, gate), 2.757),
         ("LN + modulate 36512x3072 (0.45 GB)", lambda: E.ln_modulate(x, y, B, rm, lnw, lnb, lnw, lnb, 1e-5, shift, scale), 0.0)]
if len(sys.argv) > 1 and sys.argv[1] == "attn":   # sustained comparison of attention settings: impl:emu ...
    cases = []
    for spec in sys.argv[2:]:
        impl, emu = (int(t) for t in spec.split(":"))

        def fn(impl=impl, emu=emu):
            E.attn_fwd(q, k, v, o)
        cases.append((f"self-attention impl={impl} emu={emu}", fn, 7.766, (impl, emu)))
    cases.append(("torch SDPA (library kernel, same tensors)", lambda: torch.nn.functional.scaled_dot_product_attention(q, k, v), 7.766, None))
    a8 = torch.randn(8192, 8192, device=dev).bfloat16()
    b8 = torch.randn(8192, 8192, device=dev).bfloat16()
    cases.append(("torch.matmul 8192^3 (cuBLAS; the MEASURED_PEAKS workload)", lambda: torch.matmul(a8, b8), 2 * 8192 ** 3 / 1e12, None))
else:
    cases = [c + (None,) for c in cases]
for name, fn, tflop, tune in cases:
    if tune is not None:
        E.set_tuning("attn_impl", tune[0]); E.set_tuning("attn_emu", tune[1] if tune[0] == 3 else 0)
    ms, clk, pw = sample(fn)
    extra = f"  {tflop / ms * 1e3:7.1f} TFLOP/s  {pw * ms / 1e3 / tflop:6.3f} J/TFLOP" if tflop else ""
    print(f"{name:58s} {ms:8.3f} ms  {clk:6.0f} MHz  {pw:6.0f} W  {pw * ms / 1e3:7.3f} J/call{extra}", flush=True)
