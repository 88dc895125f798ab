#!/bin/bash
mkdir -p gpurun_out
echo "=== smoke"; timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -5
echo "=== gpu tests"; timeout 900 python -m pytest tests -q -m gpu --timeout 300 2>&1 | tee gpurun_out/test_all.log | tail -15
echo "=== bench fp32"; timeout 900 python bench.py --steps 10 --warmup 3 2>&1 | tee gpurun_out/bench_fp32.log | tail -5
echo "=== bench bf16"; timeout 600 python bench.py --steps 10 --warmup 3 --precision bf16 --no-cpu-baseline 2>&1 | tee gpurun_out/bench_bf16.log | tail -3
echo "=== reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tee gpurun_out/bench_ref.log | tail -3
echo "=== ncu launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_list.log 2>&1; tail -2 gpurun_out/ncu_list.log
