#!/bin/bash
# what the driver runs at round end (smoke, gpu tests, both bench arms) plus the secondary configurations quoted in README.md
mkdir -p gpurun_out
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "=== gpu tests"; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
echo "=== reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 > gpurun_out/bench_reference.json; cut -c1-200 gpurun_out/bench_reference.json
for p in fp32 bf16 fp32_fwd; do
  extra="--no-cpu-baseline"; [ $p = fp32 ] && extra=""
  echo "=== bench $p"; timeout 900 python bench.py --precision $p $extra 2>gpurun_out/bench_$p.err | tail -1 > gpurun_out/bench_$p.json
  python -c "
import json
d=json.load(open('gpurun_out/bench_$p.json'))
print('$p value',round(d['value']),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value']),'pool frac',round(d['roofline']['frac'],3),'gemm issue',round(d['roofline_gemm']['frac_bf16_issue'],3),d.get('cpu_baseline',{}).get('value'),d['clocks'])"
done
echo "=== other configurations"
timeout 900 python tools/bench_configs.py --which 3,canon --steps 5 2>&1 | grep "^{" > gpurun_out/configs_fp32_window_off.jsonl
timeout 900 python tools/bench_configs.py --which 3,canon --steps 5 --token-window batch 2>&1 | grep "^{" > gpurun_out/configs_fp32_window_batch.jsonl
timeout 900 python tools/bench_configs.py --which canon --steps 5 --token-window batch --precision bf16 2>&1 | grep "^{" > gpurun_out/configs_bf16_window_batch.jsonl
timeout 900 python tools/bench_configs.py --which canon --steps 5 --token-window batch --precision fp32_fwd 2>&1 | grep "^{" > gpurun_out/configs_fp32fwd_window_batch.jsonl
timeout 900 python tools/bench_configs.py --which 5 --slides 500 2>&1 | grep "^{" > gpurun_out/config5_fp32.jsonl
timeout 900 python tools/bench_configs.py --which 5 --slides 500 --precision bf16 2>&1 | grep "^{" > gpurun_out/config5_bf16.jsonl
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/config*.jsonl')):
    for l in open(f):
        d=json.loads(l); print(f.split('/')[-1], {k:(round(v,2) if isinstance(v,float) else v) for k,v in d.items() if k in ('config','token_window','missing_bags_encoded_from_one_token','precision','ms_per_step','cases_per_s','e2e_slides_per_s','e2e_slides_per_s_pinned_inputs','device_resident_slides_per_s','peak_mem_gb')})
PY
