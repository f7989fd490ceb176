"""The reference's CLI contract end to end on the GPU: `python infer_cogvideo_mp_fifo.py --config <yaml>` with a tiny
random-init checkpoint tree in the diffusers layout (tools/make_tiny_checkpoint.py): from_pretrained loaders, vip.pt,
Resampler, load_video, base stage, FIFO stage, decode, and the output files the reference writes
(infer_cogvideo_mp_fifo.py:351-380: `<name>_{source,orig,fifo}_<prompt[:20]>.mp4`)."""
import glob
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _frames(path):
    import cv2
    cap = cv2.VideoCapture(path)
    n = 0
    while True:
        ok, img = cap.read()
        if not ok:
            break
        n += 1
        shape = img.shape
    cap.release()
    return n, shape


def test_cli_runs_end_to_end_on_a_tiny_checkpoint(tmp_path):
    root = str(tmp_path / "ck")
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "make_tiny_checkpoint.py"), root], check=True, cwd=ROOT)
    env = dict(os.environ, CUDA_VISIBLE_DEVICES=os.environ.get("CUDA_VISIBLE_DEVICES", "0").split(",")[0])
    r = subprocess.run([sys.executable, os.path.join(ROOT, "infer_cogvideo_mp_fifo.py"), "--config", os.path.join(root, "tiny_edit.yaml")],
                       cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    outs = glob.glob(os.path.join(root, "outputs", "tiny_*"))
    assert len(outs) == 1 and os.path.exists(os.path.join(outs[0], "config.yaml"))
    src, orig, fifo = (glob.glob(os.path.join(outs[0], f"clip1_{k}_moving gradients.mp4")) for k in ("source", "orig", "fifo"))
    assert src and orig and fifo
    assert _frames(src[0]) == (36, (96, 80, 3))      # 4 chunks x 9 frames of the conditioning clip
    assert _frames(orig[0]) == (9, (96, 80, 3))      # the base clip: 3 latent frames -> 9 frames
    assert _frames(fifo[0]) == (36, (96, 80, 3))     # 12 latent frames emitted by the FIFO stage -> 4 x 9 frames
