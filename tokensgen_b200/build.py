"""Builds libtokensgen_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

Usage: python -m tokensgen_b200.build [--force] [--verbose] [--dev]
The .so is git-ignored but travels to the GPU box with the gpurun snapshot.

--dev builds libtokensgen_b200_dev.so with -DTG_DEVELOPER: the same kernels plus the earlier attention generations and the
tg_set_tuning / tg_set_gemm_impl / tg_set_conv_impl / tg_debug_attn_trace hooks the A/B tools under tools/ use
(TG_LIB_PATH=tokensgen_b200/libtokensgen_b200_dev.so python tools/...).  The shipped library has none of them.
"""
from __future__ import annotations

import os
import subprocess
import sys
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
LIB_PATH = PKG_DIR / "libtokensgen_b200.so"
DEV_LIB_PATH = PKG_DIR / "libtokensgen_b200_dev.so"
SOURCES = ["common.cu", "gemm.cu", "attn.cu", "elementwise.cu", "conv.cu", "vae.cu"]
HEADERS = ["common.h", "ptx.cuh", "../../include/tokensgen_b200.h"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "--expt-relaxed-constexpr",
    "-Xcompiler", "-fPIC",
    "-cudart", "static",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build(dev: bool = False) -> bool:
    lib = DEV_LIB_PATH if dev else LIB_PATH
    if not lib.exists():
        return True
    t = lib.stat().st_mtime
    deps = [CSRC / s for s in SOURCES] + [(CSRC / h).resolve() for h in HEADERS] + [Path(__file__)]
    return any(d.stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False, dev: bool = False) -> Path:
    lib_path = DEV_LIB_PATH if dev else LIB_PATH
    if not force and not needs_build(dev):
        return lib_path
    objs = []
    build_dir = PKG_DIR / "build" / ("dev" if dev else "ship")
    build_dir.mkdir(parents=True, exist_ok=True)
    procs = []
    for s in SOURCES:
        obj = build_dir / (s.replace(".cu", ".o"))
        cmd = [_nvcc(), *NVCC_FLAGS, *(["-DTG_DEVELOPER"] if dev else []), "-c", str(CSRC / s), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas")
            cmd.insert(2, "-v")
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(str(obj))
    failed = False
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(f"--- nvcc {s} (rc={p.returncode})\n{out}\n")
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc compilation failed")
    link = [_nvcc(), "-shared", "-cudart", "static", "-o", str(lib_path), *objs]
    subprocess.run(link, check=True)
    return lib_path


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, dev="--dev" in sys.argv)
    print(path)
