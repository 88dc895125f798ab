#!/bin/bash
mkdir -p gpurun_out
echo "=== sanitizer memcheck"; timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_gemm.py -q -m gpu -k "gemm_nt and 128-256-64 or gated and 100 or tn_accum and 64-128" --timeout 600 > gpurun_out/memcheck_gemm.log 2>&1; grep -E "=========" gpurun_out/memcheck_gemm.log | head -40
echo "=== sanitizer memcheck kernels"; timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_kernels.py tests/test_gpu_got.py -q -m gpu -k "not 2000 and not 1000" --timeout 800 > gpurun_out/memcheck_kernels.log 2>&1; grep -E "=========" gpurun_out/memcheck_kernels.log | head -30; tail -3 gpurun_out/memcheck_kernels.log
echo "=== ncu full gemm"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 24 -c 8 -o gpurun_out/prof_gemm python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_gemm.log 2>&1; tail -1 gpurun_out/ncu_gemm.log | cut -c1-100
