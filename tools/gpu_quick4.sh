#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --timeout 300 -x 2>&1 | tail -4
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tee gpurun_out/bench_fp32.log | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value',round(d['value']),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value']) if d['e2e'] else None,'pool frac',round(d['roofline']['frac'],3),'gemm issue frac',round(d['roofline_gemm']['frac_bf16_issue'],3),'clocks',d['clocks'])
print({k:round(v,3) for k,v in d['kernel_ms_per_step'].items()})"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 330 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_list.log 2>&1
timeout 600 python tools/bench_configs.py --which 3 --precision fp32 2>&1 | tail -2 | cut -c1-420
