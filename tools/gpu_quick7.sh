#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --timeout 300 2>&1 | tail -4
timeout 900 python bench.py --no-cpu-baseline 2>&1 | tee gpurun_out/bench_fp32.json | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value',round(d['value']),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value']),'pool frac',round(d['roofline']['frac'],3),'pool ms',round(d['roofline']['avg_launch_ms'],4),'weights ms',d['roofline'].get('pool_weights_avg_launch_ms'),'gemm issue',round(d['roofline_gemm']['frac_bf16_issue'],3),d['clocks'])"
