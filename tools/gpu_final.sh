#!/bin/bash
# what the driver runs at round end, plus the profile captures committed under profiles/
mkdir -p gpurun_out
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "=== gpu tests"; timeout 900 python -m pytest tests -x -q -m gpu --timeout 300 2>&1 | tail -3
echo "=== reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 > gpurun_out/bench_reference.json; cut -c1-200 gpurun_out/bench_reference.json
echo "=== bench fp32"; timeout 900 python bench.py 2>&1 | tail -1 > gpurun_out/bench_fp32.json; python -c "
import json
d=json.load(open('gpurun_out/bench_fp32.json'))
print('value',round(d['value']),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value']),'pool frac',round(d['roofline']['frac'],3),'gemm issue',round(d['roofline_gemm']['frac_bf16_issue'],3),'cpu',round(d['cpu_baseline']['value'],1),d['clocks'])"
echo "=== bench bf16"; timeout 900 python bench.py --precision bf16 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_bf16.json; python -c "
import json
d=json.load(open('gpurun_out/bench_bf16.json'))
print('bf16 value',round(d['value']),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value']),'pool frac',round(d['roofline']['frac'],3),d['clocks'])"
echo "=== ncu launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 330 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_list.log 2>&1
echo "=== ncu full pool"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:pool_fwd_kernel -s 3 -c 1 -o gpurun_out/prof_pool_fwd python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_pool.log 2>&1
echo "=== ncu full gemm2"; timeout 900 ncu --set full --clock-control none -k regex:gemm2_kf -s 22 -c 11 -o gpurun_out/prof_gemm2 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_gemm2.log 2>&1
ls gpurun_out | head -30
