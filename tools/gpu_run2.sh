#!/bin/bash
mkdir -p gpurun_out
echo "=== got"; timeout 600 python -m pytest tests/test_gpu_got.py -q -m gpu --timeout 300 2>&1 | tee gpurun_out/test_got.log | tail -50
echo "=== all gpu tests"; timeout 900 python -m pytest tests -q -m gpu --timeout 300 2>&1 | tee gpurun_out/test_all.log | tail -40
