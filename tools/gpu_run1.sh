#!/bin/bash
# first GPU survey: each stage in its own process with its own timeout so a trap in one does not hide the rest
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "=== diag_gemm"; timeout 300 python tools/diag_gemm.py 2>&1 | tee gpurun_out/diag_gemm.log | tail -60
echo "=== kernels"; timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x --timeout 300 2>&1 | tee gpurun_out/test_kernels.log | tail -40
echo "=== gemm"; timeout 600 python -m pytest tests/test_gpu_gemm.py -q -m gpu --timeout 300 2>&1 | tee gpurun_out/test_gemm.log | tail -40
echo "=== model"; timeout 600 python -m pytest tests/test_gpu_model.py -q -m gpu --timeout 300 2>&1 | tee gpurun_out/test_model.log | tail -60
