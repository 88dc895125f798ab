#!/bin/bash
mkdir -p gpurun_out
echo "=== gpu tests"; timeout 900 python -m pytest tests -q -m gpu --timeout 300 2>&1 | tail -4
echo "=== config 3"; timeout 600 python tools/bench_configs.py --which 3 --precision fp32 2>&1 | tee gpurun_out/config3.log | tail -2 | cut -c1-800
echo "=== bench 2 gpus"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 2>&1 | tee gpurun_out/bench_2gpu.log | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('2gpu value',round(d['value']),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value']),'clocks',d['clocks'])"
echo "=== bench 1 gpu"; timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tee gpurun_out/bench_fp32.log | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('1gpu value',round(d['value']),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value']),'clocks',d['clocks'])"
echo "=== bench 1 gpu + optimizer"; timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --with-optimizer 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('1gpu+adamw value',round(d['value']),'ms',round(d['ms_per_step'],3),'launches',d['gpu_launches'])"
