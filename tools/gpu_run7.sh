#!/bin/bash
mkdir -p gpurun_out
echo "=== gpu tests"; timeout 900 python -m pytest tests -q -m gpu --timeout 300 2>&1 | tee gpurun_out/test_all.log | tail -4
echo "=== configs fp32"; timeout 900 python tools/bench_configs.py --which 3,5 --precision fp32 2>&1 | tee gpurun_out/configs_fp32.log | tail -3
echo "=== configs bf16"; timeout 900 python tools/bench_configs.py --which 3,5 --precision bf16 2>&1 | tee gpurun_out/configs_bf16.log | tail -3
echo "=== sanitizer (memcheck, small shapes)"; timeout 900 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_gemm.py -q -m gpu -k "gemm_nt and 128-256-64 or gated and 100 or tn_accum and 64-128" --timeout 600 2>&1 | tail -5
