#!/bin/bash
# `ncu --set full` of the longest kernels of the fp32-grade bench step (one launch each, second step of tools/ncu_targets.py).
# -k matches the function name without template arguments, so instances are picked by launch index: gemm2_kf_kernel is launched 11 times
# per step (L1, L2, L3, gated, attention dgrad, attention wgrad, L3 dgrad, L3 wgrad, L2 dgrad, L2 wgrad, L1 wgrad), ln_gelu_bwd 3 times.
mkdir -p gpurun_out
for spec in "gated:gemm2_kf_kernel:14" "attn_dgrad:gemm2_kf_kernel:15" "attn_wgrad:gemm2_kf_kernel:16" "gate_bwd:gate_bwd_kernel:1" "lnbwd2048:ln_gelu_bwd_kernel:3"; do
  name=${spec%%:*}; rest=${spec#*:}; regex=${rest%%:*}; skip=${rest##*:}
  timeout 300 ncu --set full --clock-control none --import-source on -k "regex:^$regex\$" --launch-skip $skip --launch-count 1 -f -o gpurun_out/r02_full_$name python tools/ncu_targets.py fp32 2 > /dev/null 2>&1
  ncu -i gpurun_out/r02_full_$name.ncu-rep --page raw --csv > gpurun_out/r02_full_${name}_raw.csv 2>/dev/null
  ncu -i gpurun_out/r02_full_$name.ncu-rep --page details --csv > gpurun_out/r02_full_${name}_details.csv 2>/dev/null
done
ls -la gpurun_out/r02_full_* | awk '{print $5, $9}'
