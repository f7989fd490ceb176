"""Regenerates every fixture of tests/golden/ from the UNMODIFIED reference (oracle/make_goldens.py, into a scratch
directory) and compares it with the committed file, bit for bit.  Build-container only (needs /root/reference); test
infrastructure like the rest of oracle/.  Exit code 0 = every committed golden is what the reference produces."""
import json
import pathlib
import sys
import tempfile

import torch

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def same(a, b) -> bool:
    import numpy as np
    if isinstance(a, np.ndarray):
        return isinstance(b, np.ndarray) and a.shape == b.shape and a.dtype == b.dtype and bool(np.array_equal(a, b))
    if torch.is_tensor(a):
        return torch.is_tensor(b) and a.shape == b.shape and a.dtype == b.dtype and torch.equal(a, b)
    if isinstance(a, dict):
        return isinstance(b, dict) and a.keys() == b.keys() and all(same(a[k], b[k]) for k in a)
    if isinstance(a, (list, tuple)):
        return isinstance(b, (list, tuple)) and len(a) == len(b) and all(same(x, y) for x, y in zip(a, b))
    return a == b


def main() -> int:
    from oracle import make_goldens as mg
    from oracle import ref_import
    committed = mg.GOLDEN
    mg.GOLDEN = pathlib.Path(tempfile.mkdtemp(prefix="goldens_"))
    import transformers  # noqa: F401  (before the stubs are on sys.path: see make_goldens._load_ref_pipeline_module)
    from transformers import AutoImageProcessor, AutoModel, T5EncoderModel, T5Tokenizer  # noqa: F401
    ref_import.enable()
    torch.set_num_threads(8)
    for fn in mg.ALL:
        fn()
    bad = []
    names = sorted(p.name for p in mg.GOLDEN.iterdir())
    if names != sorted(p.name for p in committed.iterdir()):
        bad.append("file list differs")
    for name in names:
        new, old = mg.GOLDEN / name, committed / name
        if not old.exists():
            continue
        if name.endswith(".json"):
            ok = json.load(open(new)) == json.load(open(old))
        else:
            ok = same(torch.load(new, weights_only=False), torch.load(old, weights_only=False))
        print(f"{name}: {'reproduced' if ok else 'DIFFERS'}")
        if not ok:
            bad.append(name)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
