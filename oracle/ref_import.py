"""Imports the UNMODIFIED reference (/root/reference) through oracle/stubs.  Build-container only: the GPU box has no
/root/reference, so nothing that runs there (gpu tests, smoke, bench) may call this."""
from __future__ import annotations

import os
import sys
from pathlib import Path

REFERENCE_ROOT = Path("/root/reference")
STUBS = Path(__file__).resolve().parent / "stubs"


def available() -> bool:
    return (REFERENCE_ROOT / "longvgen").is_dir()


def enable() -> None:
    if not available():
        raise RuntimeError("/root/reference is not present on this machine")
    for p in (str(STUBS), str(REFERENCE_ROOT)):
        if p not in sys.path:
            sys.path.insert(0, p)
