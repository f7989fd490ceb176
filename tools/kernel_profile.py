"""Small drivers for `ncu --set full` captures of the other hot kernels at their bench shapes.
usage: kernel_profile.py gemm_ff1 | gemm_qkv | gemm_qkv_sp | gemm_ff2 | conv | ln | norm_act"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tokensgen_b200 import _ext as E  # noqa: E402

which = sys.argv[1]
dev = "cuda"
torch.manual_seed(0)
B, F, hw, n_text, n_vip, d, H = 2, 13, 1350, 226, 480, 3072, 48
n_video = F * hw
rows = n_text + n_video + n_vip
M = B * rows
rm = E.make_rowmap(n_text, n_video, n_vip, hw, F)
table = torch.randn(B * F, 18 * d, device=dev).bfloat16()


def rep(fn, n=3):
    for _ in range(n):
        fn()
    torch.cuda.synchronize()


if which == "gemm_ff1":
    a = torch.randn(M, d, device=dev).bfloat16()
    w = (torch.randn(4 * d, d, device=dev) / d ** 0.5).bfloat16()
    bias = torch.randn(4 * d, device=dev).bfloat16()
    out = torch.empty(M, 4 * d, device=dev, dtype=torch.bfloat16)
    rep(lambda: E.gemm_bias_act(a, w, bias, out, act=E.ACT_GELU_TANH))
elif which == "gemm_ff2":
    a = torch.randn(M, 4 * d, device=dev).bfloat16()
    w = (torch.randn(d, 4 * d, device=dev) / (4 * d) ** 0.5).bfloat16()
    bias = torch.randn(d, device=dev).bfloat16()
    x = torch.randn(M, d, device=dev).bfloat16()
    gate = E.make_modvec(table[:, 5 * d:6 * d], table[:, 2 * d:3 * d], table[:, 14 * d:15 * d])
    rep(lambda: E.gemm_gate_residual(a, w, bias, x, B, rm, gate))
elif which == "gemm_qkv":
    a = torch.randn(M, d, device=dev).bfloat16()
    w = (torch.randn(6 * d, d, device=dev) / d ** 0.5).bfloat16()
    bias = torch.randn(6 * d, device=dev).bfloat16()
    n_tv = n_text + n_video
    outs = [torch.empty(B, H, n_tv if i < 3 else rows, 64, device=dev, dtype=torch.bfloat16) for i in range(6)]
    lnw, lnb = torch.ones(64, device=dev).bfloat16(), torch.zeros(64, device=dev).bfloat16()
    cos, sin = torch.rand(n_video, 64, device=dev), torch.rand(n_video, 64, device=dev)
    cosv, sinv = torch.rand(n_vip, 64, device=dev), torch.rand(n_vip, 64, device=dev)
    projs = []
    for i in range(6):
        p = E.QkvProj()
        p.out, p.out_rows = outs[i].data_ptr(), outs[i].shape[2]
        if i in (0, 1, 3, 4):
            p.ln_w, p.ln_b = lnw.data_ptr(), lnb.data_ptr()
            p.cos_video, p.sin_video = cos.data_ptr(), sin.data_ptr()
            if i >= 3:
                p.cos_vip, p.sin_vip = cosv.data_ptr(), sinv.data_ptr()
        projs.append(p)
    rep(lambda: E.qkv_rope_gemm(a, w, bias, B, H, rm, projs, 1e-6))
elif which == "gemm_qkv_sp":
    # the sequence-parallel Q/K/V GEMM of one of P = 2 ranks (rows [0, 9128) of each batch) with the head scatter going to
    # two buffers on THIS device (the kernel only sees addresses): shows what the shared-memory transposed, full-line
    # epilogue stores cost the tensor pipe
    from tokensgen_b200.seqpar import shard_rows
    P = 2
    chunk, shards = shard_rows(rows, P)
    row0, rl = shards[0]
    rms = E.make_rowmap(n_text, n_video, n_vip, hw, F, row0, rl)
    a = torch.randn(B * rl, d, device=dev).bfloat16()
    w = (torch.randn(6 * d, d, device=dev) / d ** 0.5).bfloat16()
    bias = torch.randn(6 * d, device=dev).bfloat16()
    n_tv = n_text + n_video
    bufs = [[torch.empty(B, H // P, n_tv if i < 3 else rows, 64, device=dev, dtype=torch.bfloat16) for i in range(6)] for _ in range(P)]
    lnw, lnb = torch.ones(64, device=dev).bfloat16(), torch.zeros(64, device=dev).bfloat16()
    cos, sin = torch.rand(n_video, 64, device=dev), torch.rand(n_video, 64, device=dev)
    cosv, sinv = torch.rand(n_vip, 64, device=dev), torch.rand(n_vip, 64, device=dev)
    projs = []
    for i in range(6):
        p = E.QkvProj()
        p.out_rows = n_tv if i < 3 else rows
        if i in (0, 1, 3, 4):
            p.ln_w, p.ln_b = lnw.data_ptr(), lnb.data_ptr()
            p.cos_video, p.sin_video = cos.data_ptr(), sin.data_ptr()
            if i >= 3:
                p.cos_vip, p.sin_vip = cosv.data_ptr(), sinv.data_ptr()
        projs.append(p)
    scat = E.make_qkv_scatter([[bufs[q][i].data_ptr() for q in range(P)] for i in range(6)])
    rep(lambda: E.qkv_rope_gemm(a, w, bias, B, H, rms, projs, 1e-6, scatter=scat))
elif which == "ln":
    x = torch.randn(M, d, device=dev).bfloat16()
    y = torch.empty_like(x)
    w_, b_ = torch.ones(d, device=dev).bfloat16(), torch.zeros(d, device=dev).bfloat16()
    shift = E.make_modvec(table[:, 3 * d:4 * d], table[:, 0:d], table[:, 12 * d:13 * d])
    scale = E.make_modvec(table[:, 4 * d:5 * d], table[:, d:2 * d], table[:, 13 * d:14 * d])
    rep(lambda: E.ln_modulate(x, y, B, rm, w_, b_, w_, b_, 1e-5, shift, scale))
elif which == "conv":   # decoder up-block 3 at full resolution: 8 frames 480x720, 128 -> 128, 3x3x3
    x = torch.randn(10, 480, 720, 128, device=dev).bfloat16()
    w = (torch.randn(128, 27 * 128, device=dev) / (27 * 128) ** 0.5).bfloat16()
    bias = torch.randn(128, device=dev).bfloat16()
    out = torch.empty(8, 480, 720, 128, device=dev, dtype=torch.bfloat16)
    rep(lambda: E.vae_conv(x, w, bias, 128, 3, 3, 3, 8, 480, 720, out=out))
elif which == "conv_stats":   # the same layer with the consumer GroupNorm's statistics accumulated in the epilogue (round 2)
    x = torch.randn(10, 480, 720, 128, device=dev).bfloat16()
    w = (torch.randn(128, 27 * 128, device=dev) / (27 * 128) ** 0.5).bfloat16()
    bias = torch.randn(128, device=dev).bfloat16()
    out = torch.empty(8, 480, 720, 128, device=dev, dtype=torch.bfloat16)
    stats = torch.zeros(64, device=dev, dtype=torch.float64)
    rep(lambda: E.vae_conv(x, w, bias, 128, 3, 3, 3, 8, 480, 720, out=out, stats=stats, stat_groups=32))
elif which == "norm_act":
    x = torch.randn(8, 480, 720, 128, device=dev).bfloat16()
    out = torch.empty_like(x)
    g_, b_ = torch.ones(128, device=dev).bfloat16(), torch.zeros(128, device=dev).bfloat16()

    def f():
        sums = E.vae_group_stats(x, 32)
        E.vae_norm_act(x, sums, 32, 1e-6, g_, b_, out)
    rep(f)
else:
    raise SystemExit(f"unknown kernel {which}")
