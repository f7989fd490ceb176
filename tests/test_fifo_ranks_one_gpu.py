"""A P-rank FIFO job on a ONE-GPU box: P processes share cuda:0 and talk over gloo (tools/fifo_mp_check.py, TG_CHECK_ONE_GPU=1).
Everything a multi-rank job does except the NCCL transport and the ramp sharding runs for real — base-state broadcast,
window-parallel schedule (rank r owns window r of every iteration), boundary exchange of the lookahead / write-back frames,
streaming decode with chunk c on rank c % P and the hand-off of the frames to rank 0 — and must reproduce the single-process
result bit for bit (reference: cogvideo_sampling_mp_fifo.py:230-359 runs the same windows in one process, one thread per GPU).
The NCCL + ramp-sharding variant of this check needs P GPUs: tests/test_cli_gpu.py and tools/run_r02_ngpu.sh."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(world, out, decode, port):
    env = dict(os.environ, TG_CHECK_ONE_GPU="1", TG_CHECK_DECODE="1" if decode else "0")
    tool = os.path.join(ROOT, "tools", "fifo_mp_check.py")
    if world == 1:
        cmd = [sys.executable, tool, out]
    else:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr",
               "127.0.0.1", "--master-port", str(port), tool, out]
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-12000:]
    return torch.load(out)


@pytest.mark.parametrize("decode", [False, True])
def test_two_and_three_ranks_sharing_one_gpu_reproduce_the_single_process_stage(tmp_path, decode):
    one = _run(1, str(tmp_path / "p1.pt"), decode, 0)
    assert torch.isfinite(one.float()).all()
    for world, port in ((2, 29571), (3, 29573)):
        many = _run(world, str(tmp_path / f"p{world}.pt"), decode, port + int(decode))
        assert many.shape == one.shape and torch.equal(many, one), (world, decode)
