#!/bin/bash
# N = 1, 2, 4, 8 back to back, launched as the driver launches them
mkdir -p gpurun_out
for n in 8 4 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) bench.py --gpus $n --steps 20 --warmup 5 2>&1 | tail -1 > gpurun_out/scale_$n.json
  python -c "
import json
d=json.load(open('gpurun_out/scale_$n.json'))
print('N=$n value',round(d['value']),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value']),'clocks',d['clocks'])"
done
timeout 600 python bench.py --steps 20 --warmup 5 2>&1 | tail -1 > gpurun_out/scale_1.json
python -c "
import json
d=json.load(open('gpurun_out/scale_1.json'))
print('N=1 value',round(d['value']),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value']),'cpu',d.get('cpu_baseline',{}).get('value'),'clocks',d['clocks'])"
