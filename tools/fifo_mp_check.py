"""Developer / CI tool: the tiny pipeline's FIFO stage under torchrun with P ranks (NCCL boundary exchange, base-state
broadcast) must reproduce the single-process latents bit for bit.
usage: torchrun --nproc-per-node P tools/fifo_mp_check.py out.pt   (P = 1 writes the reference)
TG_CHECK_ONE_GPU=1: every rank on cuda:0 over the gloo backend (NCCL refuses two ranks on one device) and ramp sharding off
(its sequence-parallel forward needs peer GPUs) — the window-parallel schedule, the boundary exchange, the base-state broadcast
and the streaming-decode hand-off of a P-rank job on a ONE-GPU box (tests/test_pipeline_gpu.py)."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_pipeline_gpu import _tiny_pipe  # noqa: E402
from tokensgen_b200.fifo import broadcast_base_output, cogvideo_fifo_mp_v2  # noqa: E402

world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
one_gpu = os.environ.get("TG_CHECK_ONE_GPU") == "1"
if one_gpu:
    local = 0
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    if one_gpu:
        dist.init_process_group("gloo")
        # gloo moves device tensors in collectives but not in send / recv: stage those through the host (this tool only —
        # the product's transport is NCCL)
        import torch.distributed.distributed_c10d as c10d
        _isend, _irecv, _send, _recv = c10d.isend, c10d.irecv, c10d.send, c10d.recv

        class _Landing:
            def __init__(self, work, host, target):
                self.work, self.host, self.target = work, host, target

            def wait(self):
                self.work.wait()
                self.target.copy_(self.host)
                return True

        def isend(tensor, dst=None, group=None, tag=0, group_dst=None):
            return _isend(tensor.cpu(), dst=dst, group=group, tag=tag, group_dst=group_dst)

        def irecv(tensor, src=None, group=None, tag=0, group_src=None):
            host = torch.empty(tensor.shape, dtype=tensor.dtype)
            return _Landing(_irecv(host, src=src, group=group, tag=tag, group_src=group_src), host, tensor)

        def send(tensor, dst=None, group=None, tag=0, group_dst=None):
            _send(tensor.cpu(), dst=dst, group=group, tag=tag, group_dst=group_dst)

        def recv(tensor, src=None, group=None, tag=0, group_src=None):
            host = torch.empty(tensor.shape, dtype=tensor.dtype)
            r = _recv(host, src=src, group=group, tag=tag, group_src=group_src)
            tensor.copy_(host)
            return r

        for mod in (dist, c10d):
            mod.isend, mod.irecv, mod.send, mod.recv = isend, irecv, send, recv
    else:
        dist.init_process_group("nccl", device_id=dev)
pipe = _tiny_pipe().to(dev)
decode = os.environ.get("TG_CHECK_DECODE") == "1"
base = None
if rank == 0:
    g = torch.Generator().manual_seed(3)
    frames = torch.rand(1, 36, 3, 64, 96, generator=g) * 2 - 1
    base = pipe(frames=frames, prompt_embeds=torch.randn(1, 10, 128, generator=g), negative_prompt_embeds=torch.randn(1, 10, 128, generator=g),
                height=64, width=96, num_frames_per_chunk=9, max_num_chunks=4, max_num_chunks_w_fifo=25, max_num_chunks_wo_fifo=1,
                num_inference_steps=12, guidance_scale=6.0, generator=torch.Generator().manual_seed(11), vip_scale=[0.6],
                sampling_mode="fifo", sampling_params={"num_partitions": 4}, output_type="pt" if decode else "latent", return_dict=False)
    base.condition_frames = None
else:
    pipe.preprare_for_fifo(num_inference_steps=12, vip_scale=[0.6])
base = broadcast_base_output(base, src=0, device=dev)
orig, latents, _ = cogvideo_fifo_mp_v2([pipe], base, seed=11, ramp_sharding=not one_gpu, streaming_decode=decode)
if rank == 0:                      # TG_CHECK_DECODE=1: `latents` are the decoded frames (streaming decode, chunk c on rank c % P)
    torch.save(latents.cpu(), sys.argv[1])
    print(f"world={world}: latents {tuple(latents.shape)} mean {latents.float().mean().item():.6f}")
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
