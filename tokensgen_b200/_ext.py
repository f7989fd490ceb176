"""ctypes binding of libtokensgen_b200.so (the C ABI declared in include/tokensgen_b200.h).

This is the only place Python touches the native library.  There is NO fallback: if the library is missing or a
call fails, a `TokensGenError` is raised.  PyTorch is used for device memory and streams only.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path
from typing import Optional, Sequence

import torch

_PKG = Path(__file__).resolve().parent
import os as _os
LIB_PATH = Path(_os.environ.get("TG_LIB_PATH") or (_PKG / "libtokensgen_b200.so"))  # TG_LIB_PATH: developer A/B builds


class TokensGenError(RuntimeError):
    pass


class RowMap(C.Structure):
    _fields_ = [("rows_per_batch", C.c_int), ("n_text", C.c_int), ("n_video", C.c_int), ("n_vip", C.c_int),
                ("hw", C.c_int), ("frames", C.c_int), ("row0", C.c_int), ("rows_local", C.c_int)]


class ModVec(C.Structure):
    _fields_ = [("text", C.c_void_p), ("video", C.c_void_p), ("vip", C.c_void_p),
                ("ld_text", C.c_int64), ("ld_video", C.c_int64), ("ld_vip", C.c_int64)]


class QkvProj(C.Structure):
    _fields_ = [("out", C.c_void_p), ("out_rows", C.c_int), ("ln_w", C.c_void_p), ("ln_b", C.c_void_p),
                ("cos_video", C.c_void_p), ("sin_video", C.c_void_p), ("cos_vip", C.c_void_p), ("sin_vip", C.c_void_p)]


MAX_PEERS = 8


class QkvScatter(C.Structure):
    _fields_ = [("world", C.c_int), ("peer", (C.c_void_p * MAX_PEERS) * 6)]


class AttnScatter(C.Structure):
    _fields_ = [("world", C.c_int), ("chunk", C.c_int), ("rows_per_batch", C.c_int), ("H_total", C.c_int),
                ("head0", C.c_int), ("peer", C.c_void_p * MAX_PEERS)]


class DpmStepArgs(C.Structure):
    _fields_ = [("noise_pred", C.c_void_p), ("noise_pred_f32", C.c_void_p), ("n_branches", C.c_int),
                ("guidance_scale", C.c_float),
                ("sample", C.c_void_p), ("old_x0", C.c_void_p), ("old_x0_f32", C.c_void_p),
                ("noise1", C.c_void_p), ("noise2", C.c_void_p), ("coef", C.c_void_p),
                ("prev_sample", C.c_void_p), ("x0_out", C.c_void_p), ("x0_out_f32", C.c_void_p),
                ("F", C.c_int), ("chw", C.c_int64), ("mode", C.c_int), ("guidance_scale2", C.c_float)]


class ConvArgs(C.Structure):
    _fields_ = [("x", C.c_void_p), ("T_in", C.c_int), ("H_in", C.c_int), ("W_in", C.c_int), ("Cin", C.c_int),
                ("w", C.c_void_p), ("bias", C.c_void_p), ("Cout", C.c_int), ("Cout_pad", C.c_int),
                ("kt", C.c_int), ("kh", C.c_int), ("kw", C.c_int), ("stride_hw", C.c_int), ("pad_h0", C.c_int), ("pad_w0", C.c_int),
                ("T_out", C.c_int), ("H_out", C.c_int), ("W_out", C.c_int), ("residual", C.c_void_p), ("ld_res", C.c_int64),
                ("y", C.c_void_p), ("ldy", C.c_int64), ("plane_stride", C.c_int64), ("layout", C.c_int),
                ("stats", C.c_void_p), ("stat_groups", C.c_int)]


class NormArgs(C.Structure):
    _fields_ = [("x", C.c_void_p), ("ldx", C.c_int64), ("T", C.c_int), ("H", C.c_int), ("W", C.c_int), ("C", C.c_int),
                ("groups", C.c_int), ("eps", C.c_float), ("sums", C.c_void_p), ("gamma", C.c_void_p), ("beta", C.c_void_p),
                ("zy", C.c_void_p), ("zb", C.c_void_p), ("Tz", C.c_int), ("Hz", C.c_int), ("Wz", C.c_int), ("silu", C.c_int),
                ("y", C.c_void_p), ("ldy", C.c_int64), ("ldz", C.c_int64)]


ACT_NONE, ACT_GELU_TANH, ACT_SILU = 0, 1, 2
DPM_BASE_CHAIN, DPM_BF16_CHAIN = 0, 1

# name -> (restype, argtypes): every symbol include/tokensgen_b200.h declares.
_VP, _I, _I64, _F = C.c_void_p, C.c_int, C.c_int64, C.c_float
SYMBOLS = {
    "tg_version": (C.c_int, []),
    "tg_last_error": (C.c_char_p, []),
    "tg_time_embedding": (C.c_int, [_VP, _I, _I, _I, _I, _F, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP]),
    "tg_ln_modulate": (C.c_int, [_VP, _VP, _I, _I, C.POINTER(RowMap), _VP, _VP, _VP, _VP, _F, _VP, _VP, _F,
                                 C.POINTER(ModVec), C.POINTER(ModVec), _VP]),
    "tg_gemm_bias_act": (C.c_int, [_VP, _I64, _VP, _VP, _VP, _I64, _I, _I, _I, _I, _VP]),
    "tg_gemm_gate_residual": (C.c_int, [_VP, _I64, _VP, _VP, _VP, _I64, _I, _I, _I, C.POINTER(RowMap),
                                        C.POINTER(ModVec), _VP]),
    "tg_qkv_rope_gemm": (C.c_int, [_VP, _I64, _VP, _VP, _I, _I, _I, C.POINTER(RowMap), C.POINTER(QkvProj), _I, _F, _VP]),
    "tg_qkv_rope_gemm_sp": (C.c_int, [_VP, _I64, _VP, _VP, _I, _I, _I, C.POINTER(RowMap), C.POINTER(QkvProj), _I, _F,
                                      C.POINTER(QkvScatter), _VP]),
    "tg_attn_fwd_sp": (C.c_int, [_VP, _I64, _I64, _I, _VP, _VP, _I64, _I64, _I, C.POINTER(AttnScatter), _I64, _I, _I, _F, _I,
                                 _F, _VP]),
    "tg_attn_fwd_pair_sp": (C.c_int, [_VP, _VP, _VP, _I64, _I, _I, _VP, _VP, _VP, _I64, _I64, _I, C.POINTER(AttnScatter), _I, _I,
                                      _F, _F, _VP]),
    "tg_attn_fwd": (C.c_int, [_VP, _I64, _I64, _I, _VP, _VP, _I64, _I64, _I, _VP, _I64, _I64, _I, _I, _F, _I, _F, _VP]),
    "tg_attn_fwd_pair": (C.c_int, [_VP, _VP, _VP, _I64, _I, _I, _VP, _VP, _VP, _I64, _I64, _I, _VP, _I64, _I, _I, _F, _F, _VP]),
    "tg_patchify": (C.c_int, [_VP, _VP, _I, _I, _I, _I, _I, _I, _VP]),
    "tg_unpatchify": (C.c_int, [_VP, _VP, _I, _I, _I, _I, _I, _I, _VP]),
    "tg_cfg_dpm_step": (C.c_int, [C.POINTER(DpmStepArgs), _VP]),
    "tg_queue_shift_renoise": (C.c_int, [_VP, _VP, _I, _I64, _VP, C.c_double, C.c_double, _VP]),
    "tg_vae_conv": (C.c_int, [C.POINTER(ConvArgs), _VP]),
    "tg_vae_group_stats": (C.c_int, [_VP, _I64, _I, _I64, _I, _VP, _VP]),
    "tg_vae_norm_act": (C.c_int, [C.POINTER(NormArgs), _VP]),
    "tg_vae_upsample": (C.c_int, [_VP, _VP, _I, _I, _I, _I, _I, _VP]),
    "tg_vae_avgpool_time": (C.c_int, [_VP, _VP, _I, _I64, _VP]),
    "tg_vae_to_channels_last": (C.c_int, [_VP, _VP, _I, _I, _I64, _I64, _VP]),
    "tg_vae_posterior_sample": (C.c_int, [_VP, _VP, _VP, _I64, _F, _VP]),
    "tg_vae_frames_to_rgb8": (C.c_int, [_VP, _VP, _I64, _I64, _VP]),
    "tg_vae_blend": (C.c_int, [_VP, _VP, _I64, _I, _I, _I, _I, _I, _I, _VP]),
}

_lib: Optional[C.CDLL] = None

# Instrumentation used by bench.py: number of kernel launches issued through this binding, and (when `profile` is a
# dict) CUDA-event pairs recorded around every call on the launching stream, keyed by op tag.
launch_count = 0
profile: Optional[dict] = None


_NVTX = _os.environ.get("TG_NVTX", "0") != "0"


class _Timed:
    def __init__(self, tag: str, launches: int = 1):
        self.tag, self.launches = tag, launches

    def __enter__(self):
        global launch_count
        launch_count += self.launches
        if _NVTX:   # TG_NVTX=1: one NVTX range per op tag (shows up in nsys / ncu --nvtx timelines)
            torch.cuda.nvtx.range_push(self.tag)
        if profile is not None:
            self.s = torch.cuda.Event(enable_timing=True)
            self.e = torch.cuda.Event(enable_timing=True)
            self.s.record()
        return self

    def __exit__(self, *exc):
        if _NVTX:
            torch.cuda.nvtx.range_pop()
        if profile is not None:
            self.e.record()
            profile.setdefault(self.tag, []).append((self.s, self.e))
        return False


def load() -> C.CDLL:
    """Loads the native library (once).  Raises TokensGenError when it has not been built."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise TokensGenError(
                f"{LIB_PATH} is missing: run `python -m tokensgen_b200.build` (there is no CPU/PyTorch fallback)")
        lib = C.CDLL(str(LIB_PATH))
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        if lib.tg_version() != 1:
            raise TokensGenError(f"ABI version mismatch: library reports {lib.tg_version()}")
        if hasattr(lib, "tg_set_tuning"):      # developer build only
            if _os.environ.get("TG_GEMM_IMPL"):  # A/B switch: 1 = single-CTA tiles, 2 = CTA pairs (default)
                lib.tg_set_gemm_impl(int(_os.environ["TG_GEMM_IMPL"]))
            if _os.environ.get("TG_CONV_IMPL"):
                lib.tg_set_conv_impl(int(_os.environ["TG_CONV_IMPL"]))
        _lib = lib
    return _lib


def has_tuning() -> bool:
    """True when the loaded library is a developer build (-DTG_DEVELOPER: `python -m tokensgen_b200.build --dev`,
    TG_LIB_PATH=.../libtokensgen_b200_dev.so).  The shipped library has one fixed configuration and no knobs."""
    return hasattr(load(), "tg_set_tuning")


def set_tuning(key: str, value: int) -> None:
    if not has_tuning():
        raise TokensGenError("tuning knobs exist only in the developer build (python -m tokensgen_b200.build --dev; TG_LIB_PATH)")
    lib = load()
    lib.tg_set_tuning.restype, lib.tg_set_tuning.argtypes = C.c_int, [C.c_char_p, C.c_int]
    _check(lib.tg_set_tuning(key.encode(), int(value)), "tg_set_tuning")


def _check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().tg_last_error().decode(errors="replace")
        raise TokensGenError(f"{what} failed (rc={rc}): {msg}")


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _bf16_cuda(t: torch.Tensor, name: str) -> torch.Tensor:
    if not (t.is_cuda and t.dtype == torch.bfloat16 and t.is_contiguous()):
        raise TokensGenError(f"{name}: expected a contiguous CUDA bf16 tensor, got {t.dtype} {t.device} "
                             f"contiguous={t.is_contiguous()}")
    return t


def randn_tensor(shape, generator, device, dtype) -> torch.Tensor:
    """diffusers.utils.torch_utils.randn_tensor as the reference uses it: a CPU generator draws on the CPU and the
    result is moved (so seeds reproduce across devices); a device generator draws in place."""
    gdev = generator.device if generator is not None else torch.device(device)
    if gdev.type != torch.device(device).type:
        return torch.randn(tuple(shape), generator=generator, device=gdev, dtype=dtype).to(device)
    return torch.randn(tuple(shape), generator=generator, device=device, dtype=dtype)


def make_rowmap(n_text: int, n_video: int, n_vip: int, hw: int, frames: int, row0: int = 0, rows_local: int = 0) -> RowMap:
    """row0 / rows_local: sequence-parallel shard (this rank holds rows [row0, row0+rows_local) of every batch); 0 = all."""
    return RowMap(n_text + n_video + n_vip, n_text, n_video, n_vip, hw, frames, row0, rows_local)


def rows_local(rowmap: RowMap) -> int:
    return rowmap.rows_local or rowmap.rows_per_batch


def make_modvec(text: Optional[torch.Tensor], video: Optional[torch.Tensor], vip: Optional[torch.Tensor]) -> ModVec:
    """Each argument is a 2-D bf16 view [B*frames, C] whose rows may be strided (a column slice of an AdaLN table)."""
    def ld(t):
        if t is None:
            return 0
        if t.dim() != 2 or t.stride(1) != 1 or t.dtype != torch.bfloat16 or not t.is_cuda:
            raise TokensGenError("modvec: expected 2-D CUDA bf16 with unit inner stride")
        return t.stride(0)
    return ModVec(_ptr(text), _ptr(video), _ptr(vip), ld(text), ld(video), ld(vip))


# ------------------------------------------------------------------------------------------------ ops
def time_embedding(timesteps: torch.Tensor, w1, b1, w2, b2, sincos_dim: int, flip_sin_to_cos: bool = True,
                   freq_shift: float = 0.0):
    lib = load()
    ts = timesteps.to(device=w1.device, dtype=torch.float32).contiguous()
    R, time_dim = ts.numel(), w2.shape[0]
    emb = torch.empty(R, time_dim, device=w1.device, dtype=torch.bfloat16)
    silu = torch.empty_like(emb)
    scratch = torch.empty_like(emb)
    with _Timed("time_embedding", 2):
        _check(lib.tg_time_embedding(ts.data_ptr(), R, sincos_dim, time_dim, int(flip_sin_to_cos), float(freq_shift),
                                     _bf16_cuda(w1, "w1").data_ptr(), _bf16_cuda(b1, "b1").data_ptr(),
                                     _bf16_cuda(w2, "w2").data_ptr(), _bf16_cuda(b2, "b2").data_ptr(),
                                     emb.data_ptr(), silu.data_ptr(), scratch.data_ptr(), _stream()), "tg_time_embedding")
    return emb, silu


def ln_modulate(x: torch.Tensor, out: torch.Tensor, B: int, rowmap: RowMap, ln_w, ln_b, vip_ln_w, vip_ln_b, eps: float,
                shift: ModVec, scale: ModVec, ln2_w=None, ln2_b=None, eps2: float = 1e-5) -> None:
    lib = load()
    d = x.shape[-1]
    with _Timed("ln_modulate", 1):
        _check(lib.tg_ln_modulate(_bf16_cuda(x, "x").data_ptr(), _bf16_cuda(out, "out").data_ptr(), B, d, C.byref(rowmap),
                                  _ptr(ln_w), _ptr(ln_b), _ptr(vip_ln_w), _ptr(vip_ln_b), float(eps), _ptr(ln2_w), _ptr(ln2_b),
                                  float(eps2), C.byref(shift), C.byref(scale), _stream()), "tg_ln_modulate")


def gemm_bias_act(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], out: Optional[torch.Tensor] = None,
                  act: int = ACT_NONE) -> torch.Tensor:
    """out[M,N] = act(a[M,K] @ w[N,K]^T + bias).  `a`/`out` may be row-strided 2-D views."""
    lib = load()
    if a.dim() != 2 or a.stride(1) != 1:
        raise TokensGenError("gemm_bias_act: a must be 2-D with unit inner stride")
    M, K = a.shape
    N = w.shape[0]
    if out is None:
        out = torch.empty(M, N, device=a.device, dtype=torch.bfloat16)
    if out.dim() != 2 or out.stride(1) != 1 or out.shape[0] != M or out.shape[1] != N:
        raise TokensGenError("gemm_bias_act: bad out")
    with _Timed(f"gemm_bias_act[{M}x{N}x{K}]", 1):
        _check(lib.tg_gemm_bias_act(a.data_ptr(), a.stride(0), _bf16_cuda(w, "w").data_ptr(), _ptr(bias), out.data_ptr(),
                                    out.stride(0), M, N, K, act, _stream()), "tg_gemm_bias_act")
    return out


def gemm_gate_residual(a: torch.Tensor, w: torch.Tensor, bias, x: torch.Tensor, B: int, rowmap: RowMap,
                       gate: ModVec) -> None:
    lib = load()
    M, K = a.shape
    N = w.shape[0]
    with _Timed(f"gemm_gate_residual[{M}x{N}x{K}]", 1):
        _check(lib.tg_gemm_gate_residual(_bf16_cuda(a, "a").data_ptr(), a.stride(0), _bf16_cuda(w, "w").data_ptr(), _ptr(bias),
                                         _bf16_cuda(x, "x").data_ptr(), x.stride(-2), B, N, K, C.byref(rowmap), C.byref(gate),
                                         _stream()), "tg_gemm_gate_residual")


def qkv_rope_gemm(a: torch.Tensor, w: torch.Tensor, bias, B: int, H: int, rowmap: RowMap,
                  projs: Sequence[QkvProj], ln_eps: float, scatter: Optional[QkvScatter] = None) -> None:
    """scatter: sequence-parallel peer buffers (tg_qkv_rope_gemm_sp) — head h lands on rank h // (H // world)."""
    lib = load()
    K = a.shape[-1]
    arr = (QkvProj * len(projs))(*projs)
    with _Timed(f"qkv_rope_gemm[{a.shape[0]}x{w.shape[0]}x{K}]", 1):
        if scatter is None:
            _check(lib.tg_qkv_rope_gemm(_bf16_cuda(a, "a").data_ptr(), a.stride(-2), _bf16_cuda(w, "w").data_ptr(), _ptr(bias),
                                        B, H, K, C.byref(rowmap), arr, len(projs), float(ln_eps), _stream()), "tg_qkv_rope_gemm")
        else:
            _check(lib.tg_qkv_rope_gemm_sp(_bf16_cuda(a, "a").data_ptr(), a.stride(-2), _bf16_cuda(w, "w").data_ptr(),
                                           _ptr(bias), B, H, K, C.byref(rowmap), arr, len(projs), float(ln_eps),
                                           C.byref(scatter), _stream()), "tg_qkv_rope_gemm_sp")


def make_qkv_scatter(peer_ptrs: Sequence[Sequence[int]]) -> QkvScatter:
    """peer_ptrs[p][q]: device address (peer-mapped) of projection p's [B, H/world, out_rows, 64] buffer on rank q."""
    sc = QkvScatter()
    sc.world = len(peer_ptrs[0])
    if not 1 <= sc.world <= MAX_PEERS or len(peer_ptrs) > 6:
        raise TokensGenError("make_qkv_scatter: at most 6 projections over 1..8 ranks")
    for p_, row in enumerate(peer_ptrs):
        for q_, ptr in enumerate(row):
            sc.peer[p_][q_] = ptr
    return sc


def make_attn_scatter(peer_ptrs: Sequence[int], chunk: int, rows_per_batch: int, H_total: int, head0: int) -> AttnScatter:
    """peer_ptrs[q]: device address of rank q's [B, rows_local(q), H_total*64] attention-output buffer."""
    sc = AttnScatter()
    sc.world, sc.chunk, sc.rows_per_batch, sc.H_total, sc.head0 = len(peer_ptrs), chunk, rows_per_batch, H_total, head0
    for q_, ptr in enumerate(peer_ptrs):
        sc.peer[q_] = ptr
    return sc


def attn_fwd(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, out, *, q_row0: int = 0,
             q_rows: Optional[int] = None, kv_row0: int = 0, kv_rows: Optional[int] = None, out_row0: int = 0,
             softmax_scale: Optional[float] = None, accumulate: bool = False, out_scale: float = 1.0) -> None:
    """q [B,H,Nq_alloc,64], k/v [B,H,Nkv_alloc,64] (same alloc for k and v), out [B,Nout_alloc,H*64]."""
    lib = load()
    B, H, nq_alloc, D = q.shape
    if D != 64:
        raise TokensGenError("attn_fwd: head_dim must be 64")
    nkv_alloc = k.shape[2]
    if v.shape != k.shape:
        raise TokensGenError("attn_fwd: k and v must have the same shape")
    q_rows = nq_alloc - q_row0 if q_rows is None else q_rows
    kv_rows = nkv_alloc - kv_row0 if kv_rows is None else kv_rows
    scale = D ** -0.5 if softmax_scale is None else softmax_scale
    if isinstance(out, AttnScatter):
        with _Timed(f"attn_fwd[q{q_rows},kv{kv_rows}]", 1):
            _check(lib.tg_attn_fwd_sp(_bf16_cuda(q, "q").data_ptr(), nq_alloc, q_row0, q_rows, _bf16_cuda(k, "k").data_ptr(),
                                      _bf16_cuda(v, "v").data_ptr(), nkv_alloc, kv_row0, kv_rows, C.byref(out), out_row0, B, H,
                                      float(scale), int(accumulate), float(out_scale), _stream()), "tg_attn_fwd_sp")
        return
    with _Timed(f"attn_fwd[q{q_rows},kv{kv_rows}]", 1):
        _check(lib.tg_attn_fwd(_bf16_cuda(q, "q").data_ptr(), nq_alloc, q_row0, q_rows, _bf16_cuda(k, "k").data_ptr(),
                               _bf16_cuda(v, "v").data_ptr(), nkv_alloc, kv_row0, kv_rows, _bf16_cuda(out, "out").data_ptr(),
                               out.shape[1], out_row0, B, H, float(scale), int(accumulate), float(out_scale), _stream()),
               "tg_attn_fwd")


def attn_fwd_pair(q, k, v, q_rows: int, kv_rows: int, q2, k2, v2, kv_row0_2: int, kv_rows2: int, out,
                  out_scale2: float, softmax_scale: Optional[float] = None) -> None:
    """Self-attention (q,k,v rows [0,q_rows)/[0,kv_rows)) + out_scale2 * cross-attention of q2 to rows
    [kv_row0_2, +kv_rows2) of k2/v2, one launch (tg_attn_fwd_pair)."""
    lib = load()
    B, H, alloc, D = q.shape
    alloc2 = q2.shape[2]
    scale = D ** -0.5 if softmax_scale is None else softmax_scale
    if isinstance(out, AttnScatter):
        with _Timed(f"attn_fwd_pair[q{q_rows},kv{kv_rows}+kv{kv_rows2}]", 1):
            _check(lib.tg_attn_fwd_pair_sp(_bf16_cuda(q, "q").data_ptr(), _bf16_cuda(k, "k").data_ptr(),
                                           _bf16_cuda(v, "v").data_ptr(), alloc, q_rows, kv_rows, _bf16_cuda(q2, "q2").data_ptr(),
                                           _bf16_cuda(k2, "k2").data_ptr(), _bf16_cuda(v2, "v2").data_ptr(), alloc2, kv_row0_2,
                                           kv_rows2, C.byref(out), B, H, float(scale), float(out_scale2), _stream()),
                   "tg_attn_fwd_pair_sp")
        return
    with _Timed(f"attn_fwd_pair[q{q_rows},kv{kv_rows}+kv{kv_rows2}]", 1):
        _check(lib.tg_attn_fwd_pair(_bf16_cuda(q, "q").data_ptr(), _bf16_cuda(k, "k").data_ptr(), _bf16_cuda(v, "v").data_ptr(),
                                    alloc, q_rows, kv_rows, _bf16_cuda(q2, "q2").data_ptr(), _bf16_cuda(k2, "k2").data_ptr(),
                                    _bf16_cuda(v2, "v2").data_ptr(), alloc2, kv_row0_2, kv_rows2, _bf16_cuda(out, "out").data_ptr(),
                                    out.shape[1], B, H, float(scale), float(out_scale2), _stream()), "tg_attn_fwd_pair")


def patchify(latents: torch.Tensor, p: int) -> torch.Tensor:
    lib = load()
    B, F, Cc, H, W = latents.shape
    rows = torch.empty(B * F * (H // p) * (W // p), Cc * p * p, device=latents.device, dtype=torch.bfloat16)
    with _Timed("patchify", 1):
        _check(lib.tg_patchify(_bf16_cuda(latents, "latents").data_ptr(), rows.data_ptr(), B, F, Cc, H, W, p, _stream()),
               "tg_patchify")
    return rows


def unpatchify(rows: torch.Tensor, B: int, F: int, Cc: int, H: int, W: int, p: int) -> torch.Tensor:
    lib = load()
    out = torch.empty(B, F, Cc, H, W, device=rows.device, dtype=torch.bfloat16)
    with _Timed("unpatchify", 1):
        _check(lib.tg_unpatchify(_bf16_cuda(rows, "rows").data_ptr(), out.data_ptr(), B, F, Cc, H, W, p, _stream()),
               "tg_unpatchify")
    return out


def cfg_dpm_step(noise_pred: torch.Tensor, sample: torch.Tensor, old_x0: Optional[torch.Tensor], noise1: torch.Tensor,
                 noise2: torch.Tensor, coef: torch.Tensor, guidance_scale: float, mode: int, guidance_scale2: float = 0.0):
    """noise_pred [n_branches,F,...], sample/noise [F,...] bf16; coef fp32 [F,8] (device).  Returns (prev, x0).
    Three branches (use_separate_guidance): guidance_scale / guidance_scale2 = (g_txt - 1) / (g_img - 1)."""
    lib = load()
    nb, F = noise_pred.shape[0], noise_pred.shape[1]
    chw = sample.numel() // F
    prev = torch.empty_like(sample)
    a = DpmStepArgs()
    if noise_pred.dtype == torch.float32:
        if not (noise_pred.is_cuda and noise_pred.is_contiguous()):
            raise TokensGenError("cfg_dpm_step: fp32 noise_pred must be contiguous CUDA")
        a.noise_pred_f32 = noise_pred.data_ptr()
    else:
        a.noise_pred = _bf16_cuda(noise_pred, "noise_pred").data_ptr()
    a.n_branches = nb
    a.guidance_scale = float(guidance_scale)
    a.guidance_scale2 = float(guidance_scale2)
    a.sample = _bf16_cuda(sample, "sample").data_ptr()
    a.noise1 = _bf16_cuda(noise1, "noise1").data_ptr()
    a.noise2 = _bf16_cuda(noise2, "noise2").data_ptr()
    if not (coef.is_cuda and coef.dtype == torch.float32 and coef.is_contiguous() and coef.shape == (F, 8)):
        raise TokensGenError("cfg_dpm_step: coef must be CUDA fp32 [F,8]")
    a.coef = coef.data_ptr()
    a.prev_sample = prev.data_ptr()
    a.F, a.chw, a.mode = F, chw, mode
    if mode == DPM_BF16_CHAIN:
        x0 = torch.empty_like(sample)
        a.x0_out = x0.data_ptr()
        a.old_x0 = _ptr(old_x0)
    else:
        x0 = torch.empty(sample.shape, device=sample.device, dtype=torch.float32)
        a.x0_out_f32 = x0.data_ptr()
        a.old_x0_f32 = _ptr(old_x0)
    with _Timed("cfg_dpm_step", 1):
        _check(lib.tg_cfg_dpm_step(C.byref(a), _stream()), "tg_cfg_dpm_step")
    return prev, x0


def queue_shift_renoise(queue: torch.Tensor, x0_queue: Optional[torch.Tensor], noise: torch.Tensor,
                        sqrt_one_minus_beta: float, sqrt_beta: float) -> None:
    """queue / x0_queue: bf16 [n_slots, ...] (in place); noise bf16 [...] of one slot."""
    lib = load()
    n_slots = queue.shape[0]
    chw = queue.numel() // n_slots
    with _Timed("queue_shift_renoise", 1):
        _check(lib.tg_queue_shift_renoise(_bf16_cuda(queue, "queue").data_ptr(), _ptr(x0_queue), n_slots, chw,
                                          _bf16_cuda(noise, "noise").data_ptr(), float(sqrt_one_minus_beta), float(sqrt_beta),
                                          _stream()), "tg_queue_shift_renoise")


# ------------------------------------------------------------------------------------------------ VAE ops (channels-last)
def vae_conv(x: torch.Tensor, w2d: torch.Tensor, bias: Optional[torch.Tensor], cout: int, kt: int, kh: int, kw: int,
             t_out: int, h_out: int, w_out: int, *, stride: int = 1, pad_h0: int = 1, pad_w0: int = 1,
             residual: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
             planes_out: Optional[torch.Tensor] = None, plane_stride: int = 0, stats: Optional[torch.Tensor] = None,
             stat_groups: int = 0) -> torch.Tensor:
    """x [T_in,H,W,Cin] (causal frames in front), w2d [Cout_pad, kt*kh*kw*Cin].  Returns channels-last [t_out,h_out,w_out,cout]
    (or writes channel planes into `planes_out`, a view whose first element is (n=0, t=0, h=0, w=0)).
    `stats` (zeroed fp64 [2 * stat_groups]): GroupNorm sums of the output accumulated by the epilogue (conv_stats_supported)."""
    lib = load()
    a = ConvArgs()
    T_in, H_in, W_in, Cin = x.shape
    a.x, a.T_in, a.H_in, a.W_in, a.Cin = _bf16_cuda(x, "x").data_ptr(), T_in, H_in, W_in, Cin
    a.w, a.bias, a.Cout, a.Cout_pad = _bf16_cuda(w2d, "w").data_ptr(), _ptr(bias), cout, w2d.shape[0]
    a.kt, a.kh, a.kw, a.stride_hw, a.pad_h0, a.pad_w0 = kt, kh, kw, stride, pad_h0, pad_w0
    a.T_out, a.H_out, a.W_out = t_out, h_out, w_out
    if residual is not None:
        a.residual, a.ld_res = _bf16_cuda(residual, "residual").data_ptr(), residual.shape[-1]
    if planes_out is not None:
        a.y, a.layout, a.plane_stride, a.ldy = planes_out.data_ptr(), 1, plane_stride, 0
        ret = planes_out
    else:
        if out is None:
            out = torch.empty(t_out, h_out, w_out, cout, device=x.device, dtype=torch.bfloat16)
        a.y, a.layout, a.ldy = out.data_ptr(), 0, out.stride(2)
        ret = out
    if stats is not None:
        if stats.dtype != torch.float64 or stats.numel() != 2 * stat_groups or not stats.is_cuda:
            raise TokensGenError("vae_conv: stats must be a CUDA float64 tensor of 2 * stat_groups elements")
        a.stats, a.stat_groups = stats.data_ptr(), stat_groups
    with _Timed(f"vae_conv[{t_out}x{h_out}x{w_out},{Cin}->{cout},k{kt}{kh}{kw}s{stride}]", 1):
        _check(lib.tg_vae_conv(C.byref(a), _stream()), "tg_vae_conv")
    return ret


def conv_stats_supported(cout: int, groups: int) -> bool:
    """Shapes for which tg_vae_conv can accumulate the consumer GroupNorm's statistics in its epilogue."""
    if groups <= 0 or groups > 64 or cout % 32 or cout % groups:
        return False
    cg = cout // groups
    return cg in (4, 8, 16) or cg % 32 == 0


def vae_group_stats(x: torch.Tensor, groups: int) -> torch.Tensor:
    lib = load()
    Cc = x.shape[-1]
    pixels = x.numel() // Cc
    sums = torch.zeros(2 * groups, device=x.device, dtype=torch.float64)
    with _Timed(f"vae_group_stats[{pixels}x{Cc}]", 1):
        _check(lib.tg_vae_group_stats(_bf16_cuda(x, "x").data_ptr(), pixels, Cc, Cc, groups, sums.data_ptr(), _stream()),
               "tg_vae_group_stats")
    return sums


def vae_norm_act(x: torch.Tensor, sums: torch.Tensor, groups: int, eps: float, gamma: torch.Tensor, beta: torch.Tensor,
                 out: torch.Tensor, zy: Optional[torch.Tensor] = None, zb: Optional[torch.Tensor] = None, silu: bool = True) -> None:
    """x, out: channels-last [T,H,W,C]; zy/zb: [Tz,Hz,Wz,C] or None."""
    lib = load()
    a = NormArgs()
    T, H, W, Cc = x.shape
    a.x, a.ldx, a.T, a.H, a.W, a.C, a.groups, a.eps = _bf16_cuda(x, "x").data_ptr(), Cc, T, H, W, Cc, groups, float(eps)
    a.sums, a.gamma, a.beta = sums.data_ptr(), _bf16_cuda(gamma, "gamma").data_ptr(), _bf16_cuda(beta, "beta").data_ptr()
    if zy is not None:   # [Tz,Hz,Wz,C] tables; may be column slices of a wider table (row stride = stride(2))
        for t_ in (zy, zb):
            if not (t_.is_cuda and t_.dtype == torch.bfloat16 and t_.stride(3) == 1 and t_.stride(1) == t_.shape[2] * t_.stride(2)
                    and t_.stride(0) == t_.shape[1] * t_.stride(1)):
                raise TokensGenError("vae_norm_act: zy/zb must be [Tz,Hz,Wz,C] bf16 with a uniform pixel stride")
        a.zy, a.zb = zy.data_ptr(), zb.data_ptr()
        a.Tz, a.Hz, a.Wz = zy.shape[0], zy.shape[1], zy.shape[2]
        a.ldz = zy.stride(2)
        if zb.stride(2) != zy.stride(2):
            raise TokensGenError("vae_norm_act: zy and zb must share the pixel stride")
    a.silu = int(silu)
    if not (out.is_cuda and out.dtype == torch.bfloat16 and out.stride(-1) == 1):
        raise TokensGenError("vae_norm_act: bad out")
    a.y, a.ldy = out.data_ptr(), out.stride(2)
    with _Timed(f"vae_norm_act[{T * H * W}x{Cc}]", 1):
        _check(lib.tg_vae_norm_act(C.byref(a), _stream()), "tg_vae_norm_act")


def vae_upsample(x: torch.Tensor, compress_time: bool) -> torch.Tensor:
    lib = load()
    T, H, W, Cc = x.shape
    T2 = ((2 * T - 1) if T % 2 == 1 else 2 * T) if (compress_time and T > 1) else T
    y = torch.empty(T2, 2 * H, 2 * W, Cc, device=x.device, dtype=torch.bfloat16)
    with _Timed("vae_upsample", 1):
        _check(lib.tg_vae_upsample(_bf16_cuda(x, "x").data_ptr(), y.data_ptr(), T, H, W, Cc, int(compress_time), _stream()),
               "tg_vae_upsample")
    return y


def vae_avgpool_time(x: torch.Tensor) -> torch.Tensor:
    lib = load()
    T = x.shape[0]
    T2 = 1 + (T - 1) // 2 if T % 2 == 1 else T // 2
    y = torch.empty((T2,) + tuple(x.shape[1:]), device=x.device, dtype=torch.bfloat16)
    with _Timed("vae_avgpool_time", 1):
        _check(lib.tg_vae_avgpool_time(_bf16_cuda(x, "x").data_ptr(), y.data_ptr(), T, x[0].numel(), _stream()),
               "tg_vae_avgpool_time")
    return y


def vae_to_channels_last(x: torch.Tensor, cpad: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x [C,T,H,W] (a view with contiguous [T,H,W] planes) -> [T,H,W,cpad]."""
    lib = load()
    Cc, T, H, W = x.shape
    if not (x.is_cuda and x.dtype == torch.bfloat16 and x[0].is_contiguous()):
        raise TokensGenError("vae_to_channels_last: expected CUDA bf16 [C,T,H,W] with contiguous planes")
    if out is None:
        out = torch.empty(T, H, W, cpad, device=x.device, dtype=torch.bfloat16)
    with _Timed("vae_to_channels_last", 1):
        _check(lib.tg_vae_to_channels_last(x.data_ptr(), _bf16_cuda(out, "out").data_ptr(), Cc, cpad, T * H * W, x.stride(0), _stream()),
               "tg_vae_to_channels_last")
    return out


def vae_posterior_sample(moments: torch.Tensor, eps: torch.Tensor, scale: float) -> torch.Tensor:
    """moments [2L, ...] planes (mean | logvar), eps [L, ...] -> z [L, ...] = sample * scale."""
    lib = load()
    z = torch.empty_like(eps)
    with _Timed("vae_posterior_sample", 1):
        _check(lib.tg_vae_posterior_sample(_bf16_cuda(moments, "moments").data_ptr(), _bf16_cuda(eps, "eps").data_ptr(), z.data_ptr(),
                                           eps.numel(), float(scale), _stream()), "tg_vae_posterior_sample")
    return z


def vae_blend(a: torch.Tensor, b: torch.Tensor, extent: int, axis: int) -> None:
    """a, b: [planes..., H, W] contiguous bf16; blends the leading `extent` rows (axis 0) / columns (axis 1) of b in place."""
    lib = load()
    Ha, Wa, Hb, Wb = a.shape[-2], a.shape[-1], b.shape[-2], b.shape[-1]
    planes = b.numel() // (Hb * Wb)
    with _Timed("vae_blend", 1):
        _check(lib.tg_vae_blend(_bf16_cuda(a, "a").data_ptr(), _bf16_cuda(b, "b").data_ptr(), planes, Ha, Wa, Hb, Wb, extent, axis,
                                _stream()), "tg_vae_blend")


def vae_frames_to_rgb8(video: torch.Tensor) -> torch.Tensor:
    """video [3, F, H, W] bf16 in [-1, 1] (contiguous planes) -> [F, H, W, 3] uint8 (tg_vae_frames_to_rgb8)."""
    lib = load()
    if video.dim() != 4 or video.shape[0] != 3 or not (video.is_cuda and video.dtype == torch.bfloat16 and video[0].is_contiguous()):
        raise TokensGenError("vae_frames_to_rgb8: expected CUDA bf16 [3, F, H, W] with contiguous planes")
    _, F, H, W = video.shape
    out = torch.empty(F, H, W, 3, device=video.device, dtype=torch.uint8)
    with _Timed("vae_frames_to_rgb8", 1):
        _check(lib.tg_vae_frames_to_rgb8(video.data_ptr(), out.data_ptr(), F * H * W, video.stride(0), _stream()),
               "tg_vae_frames_to_rgb8")
    return out
