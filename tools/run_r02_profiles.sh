#!/bin/bash
# Round-2 evidence run on ONE B200 (numbers under a profiler are never bench values): the default bench line, the ncu launch
# list of the same command, `--set full` captures of the shipped attention launch and of the conv kernel with epilogue
# statistics, the config-5 VAE sweep, and the measured 1-GPU FIFO stage that bench.py's fifo_stage speed-ups refer to.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r02.log 2> gpurun_out/bench_r02.err; echo "bench rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"attn|gemm|ln_modulate|patch_map|dpm|time_embed|queue_shift|conv|norm_act|group_stats" -c 700 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-eager-baseline --no-vae > gpurun_out/ncu_bench.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn3 -s 2 -c 1 -f -o gpurun_out/attn_pair_full python tools/attn_profile.py pair > gpurun_out/ncu_attn_pair.log 2>&1; echo "ncu attn pair rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv3_kernel -s 2 -c 1 -f -o gpurun_out/conv_stats_full python tools/kernel_profile.py conv_stats > gpurun_out/ncu_conv_stats.log 2>&1; echo "ncu conv rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:norm_act_staged -s 5 -c 1 -f -o gpurun_out/norm_act_staged_full python tools/norm_act_ab.py > gpurun_out/ncu_norm.log 2>&1; echo "ncu norm rc=$?"
timeout 120 python tools/norm_act_ab.py > gpurun_out/norm_act_ab.txt 2>&1
timeout 600 python tools/vae_bench.py 13 25 49 97 193 385 > gpurun_out/vae_sweep_r02.jsonl 2> gpurun_out/vae_sweep_r02.err; echo "vae sweep rc=$?"
if [ -z "$SKIP_FIFO" ]; then
timeout 900 python tools/fifo_full_run.py ${FIFO_CHUNKS:-3} > gpurun_out/fifo_stage_p1_r02.json 2> gpurun_out/fifo_stage_p1_r02.err; echo "fifo 1-GPU rc=$?"
fi
tail -c 1500 gpurun_out/bench_r02.log; tail -2 gpurun_out/fifo_stage_p1_r02.json
