mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/smi.txt 2>&1
timeout 700 python tools/gpu_check.py > gpurun_out/gpu_check.log 2>&1; echo "gpu_check rc=$?" 
timeout 900 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/bench.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"attn|gemm|ln_modulate|patch_map|dpm|time_embed|queue_shift|conv|groupnorm|vae" -c 700 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-eager-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn3 -c 1 -f -o gpurun_out/attn3_full python tools/attn_profile.py 48 > gpurun_out/ncu_attn.log 2>&1; echo "ncu attn rc=$?"
tail -n 5 gpurun_out/gpu_check.log; tail -n 5 gpurun_out/pytest_gpu.log
