#!/bin/bash
# Round-2 multi-GPU validation on an N-GPU box (N = $1): FIFO stage P = 1 vs P = N on the tiny pipeline (ramp sharding with
# groups of N / N/2 / ... ranks + boundary exchange: bit-identical latents) and bench.py at N ranks with the real FIFO stage.
N=${1:-4}
mkdir -p gpurun_out
CUDA_VISIBLE_DEVICES=0 python tools/fifo_mp_check.py /tmp/fifo_p1.pt 2>&1 | tail -1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561 tools/fifo_mp_check.py /tmp/fifo_pN.pt 2>&1 | tail -2
python -c "
import torch
a, b = torch.load('/tmp/fifo_p1.pt'), torch.load('/tmp/fifo_pN.pt')
print('FIFO stage P=1 vs P=$N (ramp sharding on):', 'bit-identical' if torch.equal(a, b) else 'DIFFERENT %g' % (a.float()-b.float()).abs().max().item())
" | tee gpurun_out/fifo_p1_vs_p$N.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus $N --steps 5 --warmup 3 --fifo-chunks ${FIFO_CHUNKS:-2} > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err; echo "bench N=$N rc=$?"
tail -c 2200 gpurun_out/bench_n$N.log; tail -3 gpurun_out/bench_n$N.err
