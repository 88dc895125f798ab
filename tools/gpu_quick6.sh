#!/bin/bash
mkdir -p gpurun_out
echo "=== diag"; timeout 300 python tools/diag_gemm.py 2>&1 | grep -E "FAIL|TN rand|Traceback|Error|timeout" | head -20
echo "=== tests"; timeout 900 python -m pytest tests -q -m gpu --timeout 300 2>&1 | tail -6
echo "=== bench"
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tee gpurun_out/bench_fp32.log | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value',round(d['value']),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value']) if d['e2e'] else None,'pool frac',round(d['roofline']['frac'],3),'gemm issue frac',round(d['roofline_gemm']['frac_bf16_issue'],3),'clocks',d['clocks'])
print({k:round(v,3) for k,v in d['kernel_ms_per_step'].items()})"
