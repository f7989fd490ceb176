"""The committed fixtures are what the UNMODIFIED reference produces: regenerate them here (build container only — the
reference tree does not travel to the GPU box) and compare bit for bit."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not os.path.isdir("/root/reference/longvgen"), reason="the reference tree is not on this machine")
def test_committed_goldens_reproduce_from_the_unmodified_reference():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "check_goldens.py")], cwd=ROOT, capture_output=True,
                       text=True, timeout=1200)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("reproduced") == 10 and "DIFFERS" not in r.stdout
