"""The reference's CLI contract end to end on the GPU: `python infer_cogvideo_mp_fifo.py --config <yaml>` with a tiny
random-init checkpoint tree in the diffusers layout (tools/make_tiny_checkpoint.py): from_pretrained loaders, vip.pt,
Resampler, load_video, base stage, FIFO stage, decode, and the output files the reference writes
(infer_cogvideo_mp_fifo.py:351-380: `<name>_{source,orig,fifo}_<prompt[:20]>.mp4`)."""
import glob
import os
import subprocess
import sys

import pytest
import torch

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
    # a second process writes the same videos: every draw (initial noise, conditioning-clip posterior, FIFO re-noise) comes
    # from generators seeded by the yaml, none from the device's unseeded global RNG
    import numpy as np
    first = {k: _all_frames(v[0]) for k, v in (("orig", orig), ("fifo", fifo))}
    import shutil
    shutil.rmtree(outs[0])
    r = subprocess.run([sys.executable, os.path.join(ROOT, "infer_cogvideo_mp_fifo.py"), "--config", os.path.join(root, "tiny_edit.yaml")],
                       cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    outs = glob.glob(os.path.join(root, "outputs", "tiny_*"))
    for k in ("orig", "fifo"):
        again = _all_frames(glob.glob(os.path.join(outs[0], f"clip1_{k}_moving gradients.mp4"))[0])
        assert np.array_equal(first[k], again), k


def _all_frames(path):
    import cv2
    import numpy as np
    cap = cv2.VideoCapture(path)
    out = []
    while True:
        ok, img = cap.read()
        if not ok:
            break
        out.append(img)
    cap.release()
    return np.stack(out)


def test_cli_gen_flow_t2to_then_to2v(tmp_path):
    """config/infer/gen.yaml's flow (use_2nd_stage): T2To tokens transformer -> PCA un-projection -> condensed tokens ->
    To2V base clip -> FIFO -> decode, on the tiny checkpoint tree."""
    root = str(tmp_path / "ck")
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "make_tiny_checkpoint.py"), root], check=True, cwd=ROOT)
    env = dict(os.environ, CUDA_VISIBLE_DEVICES=os.environ.get("CUDA_VISIBLE_DEVICES", "0").split(",")[0])
    r = subprocess.run([sys.executable, os.path.join(ROOT, "infer_cogvideo_mp_fifo.py"), "--config", os.path.join(root, "tiny_gen.yaml")],
                       cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    outs = glob.glob(os.path.join(root, "outputs", "tinygen_*"))
    assert len(outs) == 1
    emb, orig, fifo = (glob.glob(os.path.join(outs[0], f"clip1_{k}_moving gradients.*")) for k in ("embeds", "orig", "fifo"))
    assert emb and orig and fifo and not glob.glob(os.path.join(outs[0], "clip1_source_*"))
    e = torch.load(emb[0], weights_only=True)
    assert tuple(e.shape) == (8, 256, 2, 3) and torch.isfinite(e.float()).all()   # 4 chunks x 2 temporal queries, 2 x 3 grid
    assert _frames(orig[0]) == (9, (96, 80, 3))
    assert _frames(fifo[0]) == (36, (96, 80, 3))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("mode", ["cfg_parallel", "sequence_parallel", "gen_sequence_parallel"])
def test_cli_two_ranks_writes_the_same_videos_as_one_rank(tmp_path, mode):
    """1 rank vs torchrun x2 — base clip CFG-parallel (one guidance branch per rank) or sequence-parallel (every DiT
    forward sharded over both ranks, conditioning chunks encoded one per rank), FIFO stage window-parallel, decode
    clip-parallel: identical mp4s."""
    import numpy as np
    roots = [str(tmp_path / "one"), str(tmp_path / "two")]
    for r_ in roots:
        subprocess.run([sys.executable, os.path.join(ROOT, "tools", "make_tiny_checkpoint.py"), r_], check=True, cwd=ROOT)
    yaml_name = "tiny_gen.yaml" if mode.startswith("gen") else "tiny_edit.yaml"
    if mode.endswith("sequence_parallel"):
        with open(os.path.join(roots[1], yaml_name), "a") as f:
            f.write("sequence_parallel: true\n")
    cli = os.path.join(ROOT, "infer_cogvideo_mp_fifo.py")
    env1 = dict(os.environ, CUDA_VISIBLE_DEVICES=os.environ.get("CUDA_VISIBLE_DEVICES", "0").split(",")[0])
    r = subprocess.run([sys.executable, cli, "--config", os.path.join(roots[0], yaml_name)], cwd=ROOT, env=env1,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    if mode == "cfg_parallel":
        # the REFERENCE's invocation: one plain `python` command, GPUs from CUDA_VISIBLE_DEVICES (its :191,:384-389) — the
        # CLI starts one rank per visible GPU itself (launch_plan)
        two = ",".join((os.environ.get("CUDA_VISIBLE_DEVICES") or "0,1").split(",")[:2])
        env2 = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
        r = subprocess.run([sys.executable, cli, "--config", os.path.join(roots[1], yaml_name)], cwd=ROOT,
                           env=dict(env2, CUDA_VISIBLE_DEVICES=two), capture_output=True, text=True, timeout=600)
        assert "Running on gpus: [0, 1]" in r.stdout
    else:
        r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                            "127.0.0.1", "--master-port", "29547", cli, "--config", os.path.join(roots[1], yaml_name)],
                           cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    for kind in ("orig", "fifo"):
        a, b = (_all_frames(glob.glob(os.path.join(r_, "outputs", "tiny*", f"clip1_{kind}_*.mp4"))[0]) for r_ in roots)
        assert a.shape == b.shape and np.array_equal(a, b), kind
