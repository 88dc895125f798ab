#!/bin/bash
mkdir -p gpurun_out
echo "=== gpu tests"; timeout 900 python -m pytest tests -q -m gpu --timeout 300 -x 2>&1 | tee gpurun_out/test_all.log | tail -15
echo "=== bench fp32"; timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tee gpurun_out/bench_fp32.log | tail -2
echo "=== bench bf16"; timeout 600 python bench.py --steps 10 --warmup 3 --precision bf16 --no-cpu-baseline 2>&1 | tee gpurun_out/bench_bf16.log | tail -2
echo "=== ncu launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 330 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_list.log 2>&1; tail -1 gpurun_out/ncu_list.log | cut -c1-200
