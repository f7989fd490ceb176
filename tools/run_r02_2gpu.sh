#!/bin/bash
# Round-2 multi-GPU validation on a 2-GPU box: the 2-rank tests the 1-GPU driver run skips, FIFO stage P = 1 vs P = 2
# (ramp sharding + boundary exchange: bit-identical latents), and bench.py at N = 2 with the real FIFO stage.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -rs 2>&1 | tail -15 > gpurun_out/pytest_2gpu.log; tail -8 gpurun_out/pytest_2gpu.log
CUDA_VISIBLE_DEVICES=0 python tools/fifo_mp_check.py /tmp/fifo_p1.pt 2>&1 | tail -1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 tools/fifo_mp_check.py /tmp/fifo_p2.pt 2>&1 | tail -2
python -c "
import torch
a, b = torch.load('/tmp/fifo_p1.pt'), torch.load('/tmp/fifo_p2.pt')
print('FIFO stage P=1 vs P=2 (ramp sharding on):', 'bit-identical' if torch.equal(a, b) else 'DIFFERENT %g' % (a.float()-b.float()).abs().max().item())
" | tee gpurun_out/fifo_p1_vs_p2.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 2 --steps 5 --warmup 3 --fifo-chunks ${FIFO_CHUNKS:-1} > gpurun_out/bench_n2.log 2> gpurun_out/bench_n2.err; echo "bench N=2 rc=$?"
tail -c 2500 gpurun_out/bench_n2.log; tail -3 gpurun_out/bench_n2.err
