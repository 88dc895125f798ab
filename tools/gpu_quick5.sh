#!/bin/bash
mkdir -p gpurun_out
echo "=== diag (2cta default)"; timeout 300 python tools/diag_gemm.py 2>&1 | grep -E "FAIL|NT rand.*nsplit=3|Traceback|Error|timeout" | head -20
echo "=== tests"; timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_model.py tests/test_gpu_fullsize.py -q -m gpu --timeout 300 2>&1 | tail -6
echo "=== bounds"; timeout 300 python tools/gemm_bounds.py 2>&1 | tail -4 | cut -c1-120
for f in 1 0; do
echo "=== bench 2cta=$f"
MDL_GEMM_2CTA=$f timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tee gpurun_out/bench_fp32_2cta$f.log | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value',round(d['value']),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value']) if d['e2e'] else None,'pool frac',round(d['roofline']['frac'],3),'gemm issue frac',round(d['roofline_gemm']['frac_bf16_issue'],3),'clocks',d['clocks'])
print({k:round(v,3) for k,v in d['kernel_ms_per_step'].items()})"
done
