# ncu --set full captures of the non-attention hot kernels at their bench shapes (one launch each) + the VAE sweep
mkdir -p gpurun_out
for k in gemm_ff1:gemm2 gemm_qkv:gemm2 gemm_ff2:gemm2 conv:conv3_kernel ln:ln_modulate norm_act:norm_act; do
  name=${k%%:*}; pat=${k##*:}
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$pat -s 2 -c 1 -f -o gpurun_out/${name}_full python tools/kernel_profile.py $name > gpurun_out/ncu_${name}.log 2>&1; echo "$name rc=$?"
done
timeout 300 ncu --set full --clock-control none -k regex:group_stats -s 2 -c 1 -f -o gpurun_out/group_stats_full python tools/kernel_profile.py norm_act > /dev/null 2>&1; echo "group_stats rc=$?"
timeout 400 python tools/vae_bench.py 13 > gpurun_out/vae_bench.log 2>&1; tail -4 gpurun_out/vae_bench.log | cut -c1-700
