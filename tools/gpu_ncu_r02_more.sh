mkdir -p gpurun_out
for spec in "lnbwd512:ln_gelu_bwd_kernel:4" "lnfwd512:ln_gelu_fwd_kernel:3" "lnfwd2048:ln_gelu_fwd_kernel:5" "pool_bwd:pool_bwd_dlogit_kernel:1" "lnbwd2048_v2:ln_gelu_bwd_kernel:3"; do
  name=${spec%%:*}; rest=${spec#*:}; regex=${rest%%:*}; skip=${rest##*:}
  timeout 300 ncu --set full --clock-control none --import-source on -k "regex:^$regex\$" --launch-skip $skip --launch-count 1 -f -o gpurun_out/r02_full_$name python tools/ncu_targets.py fp32 2 > /dev/null 2>&1
done
ls -la gpurun_out/r02_full_*.ncu-rep | awk '{print $5, $9}'
timeout 200 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "ln_gelu" 2>&1 | tail -2
