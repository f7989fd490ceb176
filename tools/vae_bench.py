"""BASELINE.json configs[4]: 3D causal VAE decode throughput sweep, T_lat latent frames of 60x90 (-> 480x720), full-size
CogVideoX VAE (random-init weights of the true shapes), one B200.  Also encode of one 49-frame clip.
Prints one JSON line per point: ms, TFLOP/s of the conv kernels (algorithmic conv FLOPs / total time) and the per-op
breakdown, against MEASURED_PEAKS.json.
usage: python tools/vae_bench.py [T_lat ...]   (default 13 25 49)"""
import json
import os
import re
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tokensgen_b200 import _ext as E  # noqa: E402
from tokensgen_b200.vae import AutoencoderKLCogVideoX  # noqa: E402


def conv_flops(tag: str) -> float:
    m = re.match(r"vae_conv\[(\d+)x(\d+)x(\d+),(\d+)->(\d+),k(\d)(\d)(\d)s(\d)\]", tag)
    t, h, w, cin, cout, kt, kh, kw, _ = (int(x) for x in m.groups())
    return 2.0 * t * h * w * cin * cout * kt * kh * kw


def run(fn, tiled):
    torch.cuda.synchronize()
    fn()  # warm-up (allocations, tensor maps)
    torch.cuda.synchronize()
    E.profile = {}
    s, e = torch.cuda.Event(True), torch.cuda.Event(True)
    s.record()
    fn()
    e.record()
    torch.cuda.synchronize()
    prof, E.profile = E.profile, None
    ms = s.elapsed_time(e)
    per = {}
    flops = 0.0
    hbm = {"vae_norm_act": [0.0, 0.0], "vae_group_stats": [0.0, 0.0]}   # [algorithmic bytes, ms]
    for k, v in prof.items():
        t = sum(a.elapsed_time(b) for a, b in v)
        if k.startswith("vae_conv"):
            flops += conv_flops(k) * len(v)
            per["vae_conv"] = per.get("vae_conv", 0.0) + t
        elif k.startswith(("vae_norm_act[", "vae_group_stats[")):
            name = k.split("[")[0]
            px, c = (int(x) for x in k.split("[")[1].rstrip("]").split("x"))
            hbm[name][0] += (2 if name == "vae_norm_act" else 1) * px * c * 2.0 * len(v)   # read (+ write) of the bf16 activation
            hbm[name][1] += t
            per[name] = per.get(name, 0.0) + t
        elif k.startswith("gemm_bias_act"):
            per["gemm_1x1"] = per.get("gemm_1x1", 0.0) + t
        else:
            per[k] = per.get(k, 0.0) + t
    if os.environ.get("TG_VAE_BY_LAYER") == "1":      # developer view: every distinct launch shape, worst total first
        rows = []
        for k, v in prof.items():
            t = sum(a.elapsed_time(b) for a, b in v)
            tf = conv_flops(k) * len(v) / t / 1e9 if k.startswith("vae_conv") else 0.0
            rows.append((t, k, len(v), tf))
        for t, k, n, tf in sorted(rows, reverse=True)[:40]:
            print(f"  {t:8.2f} ms  x{n:<4d} {k}" + (f"  {tf:7.1f} TFLOP/s" if tf else ""), file=sys.stderr)
    out = {k: round(v, 2) for k, v in sorted(per.items(), key=lambda kv: -kv[1])}
    for name, (b, t) in hbm.items():
        if t > 0:
            out[name + "_GBps"] = round(b / t / 1e6, 0)
    return ms, flops, out


def build_vae(seed: int = 0):
    torch.manual_seed(seed)
    with torch.device("meta"):
        vae = AutoencoderKLCogVideoX(scaling_factor=0.7)
    vae = vae.to_empty(device="cuda").to(torch.bfloat16)
    g = torch.Generator(device="cuda").manual_seed(seed)
    with torch.no_grad():
        for name, p in vae.named_parameters():
            if p.dim() >= 2:
                p.normal_(0.0, 0.02, generator=g)
            elif name.endswith("weight"):
                p.fill_(1.0)
            else:
                p.normal_(0.0, 0.02, generator=g)
    return vae.eval()


def point(vae, op: str, T: int, tiled: bool, peaks: dict, g=None):
    """One measured point: op = "decode" (T latent frames of 60 x 90, ONE causal stream: frame batches 3, 2, 2, ... with the
    conv cache carried, autoencoder_kl_cogvideox.py:1144-1157) or "encode" (T pixel frames of 480 x 720)."""
    g = g or torch.Generator().manual_seed(42)
    tf_peak, hbm_peak = peaks.get("bf16_tflops_sustained", 1400.0), peaks.get("hbm_gbs", 6650.0)
    vae.enable_tiling() if tiled else vae.disable_tiling()
    with torch.no_grad():
        if op == "decode":
            z = torch.randn(1, 16, T, 60, 90, generator=g).cuda().bfloat16()
            ms, fl, per = run(lambda: vae.decode(z).sample, tiled)
            px = (T - 1) * 4 + 1
        else:
            x = (torch.rand(1, 3, T, 480, 720, generator=g) * 2 - 1).cuda().bfloat16()
            ms, fl, per = run(lambda: vae.encode(x).latent_dist.parameters, tiled)
            px = T
    vae.disable_tiling()
    out = {"op": op, "tiled": tiled, "latent_frames": T if op == "decode" else (T - 1) // 4 + 1, "pixel_frames": px, "ms": round(ms, 1),
           "pixel_frames_per_s": round(px / ms * 1e3, 1), "conv_tflop": round(fl / 1e12, 1),
           "tflops_whole_pass": round(fl / ms / 1e9, 1), "frac_of_sustained_peak": round(fl / ms / 1e9 / tf_peak, 3),
           "conv_kernel_tflops": round(fl / per.get("vae_conv", ms) / 1e9, 1),
           "conv_kernel_frac_of_sustained_peak": round(fl / per.get("vae_conv", ms) / 1e9 / tf_peak, 3), "ms_by_op": per}
    from tokensgen_b200 import vae as V
    if tiled and V._TILE_STREAMS > 1:
        # the tiles run on several CUDA streams: the per-op event intervals overlap each other, only the pass time is meaningful
        out["ms_by_op"] = {"note": f"tiles on {V._TILE_STREAMS} streams: per-op intervals overlap (TG_VAE_TILE_STREAMS=1 for a breakdown)"}
        del out["conv_kernel_tflops"], out["conv_kernel_frac_of_sustained_peak"]
        return out
    for name in ("vae_norm_act", "vae_group_stats"):
        if name + "_GBps" in per:
            out[name + "_frac_of_hbm_peak"] = round(per[name + "_GBps"] / hbm_peak, 3)
    return out


def main():
    frames = [int(a) for a in sys.argv[1:]] or [13, 25, 49]
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    vae = build_vae()
    for T in frames:
        print(json.dumps(point(vae, "decode", T, False, peaks)), flush=True)
    print(json.dumps(point(vae, "decode", 13, True, peaks)), flush=True)
    print(json.dumps(point(vae, "encode", 49, False, peaks)), flush=True)
    print(json.dumps(point(vae, "encode", 49, True, peaks)), flush=True)


if __name__ == "__main__":
    main()
