"""Recipe for `baseline/_ref/`: the UNMODIFIED reference files of the hot path, copied from where they lie under
/root/reference, next to the import stubs that let them run without diffusers / xformers / accelerate (oracle/stubs).

    python -m oracle.vendor_reference            (also run by __graft_entry__.build() when /root/reference is present)

`baseline/_ref/` is git-ignored (no reference source enters the history) but NOT gpurun-ignored, so it travels to the GPU box,
where /root/reference does not exist: there `bench.py --impl reference`, `cpu_baseline` and `gpu_eager_baseline` time the
reference's own `CogVideoXBlock` (+ `set_vip_layers`) from this directory.  Test infrastructure, like the rest of oracle/:
nothing on the product path imports it."""
from __future__ import annotations

import shutil
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
REFERENCE = Path("/root/reference")
DEST = ROOT / "baseline" / "_ref"
# SURVEY §2.1 ★ files the block / model / VAE / scheduler / resampler forward needs (cogvideox_transformer_3d.py:221-332 etc.)
FILES = [
    "longvgen/models/__init__.py",
    "longvgen/models/cogvideox_transformer_3d.py",
    "longvgen/models/attention_processor.py",
    "longvgen/models/normalization.py",
    "longvgen/models/embeddings.py",
    "longvgen/models/autoencoder_kl_cogvideox.py",
    "longvgen/schedulers/__init__.py",
    "longvgen/schedulers/scheduling_dpm_cogvideox.py",
    "longvgen/schedulers/scheduling_ddim_cogvideox.py",
    "longvgen/video_ipadapter/__init__.py",
    "longvgen/video_ipadapter/resampler.py",
]


def vendor(verbose: bool = True) -> bool:
    if not (REFERENCE / "longvgen").is_dir():
        if verbose:
            print("oracle.vendor_reference: /root/reference is not on this machine; keeping", DEST if DEST.exists() else "nothing")
        return DEST.exists()
    if DEST.exists():
        shutil.rmtree(DEST)
    for rel in FILES:
        dst = DEST / rel
        dst.parent.mkdir(parents=True, exist_ok=True)
        shutil.copyfile(REFERENCE / rel, dst)
    shutil.copytree(ROOT / "oracle" / "stubs", DEST / "stubs", ignore=shutil.ignore_patterns("__pycache__"))
    (DEST / "README").write_text("Unmodified files of Vicky0522/TokensGen copied by oracle/vendor_reference.py (git-ignored) + oracle/stubs.\n")
    if verbose:
        print(f"oracle.vendor_reference: {len(FILES)} reference files + stubs -> {DEST}")
    return True


def enable() -> bool:
    """Puts baseline/_ref (and its stubs) FIRST on sys.path so `import longvgen.models...` resolves to the reference's files.
    Returns False when the directory has not been built."""
    if not (DEST / "longvgen" / "models" / "cogvideox_transformer_3d.py").exists():
        return False
    for p in (str(DEST / "stubs"), str(DEST)):
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
    stale = [m for m in sys.modules if m == "longvgen" or m.startswith("longvgen.")]
    for m in stale:                      # an earlier import of the alias package must not shadow the reference's modules
        del sys.modules[m]
    return True


if __name__ == "__main__":
    sys.exit(0 if vendor() else 1)
