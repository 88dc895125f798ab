#!/bin/bash
# ncu evidence for the round-2 kernels (one B200): (1) per-launch durations of one bench step, (2) tensor-pipe / DRAM metrics of
# every kernel of a step, (3) one --set full capture of the attention-pooling kernel (the bench line's roofline.traffic).
mkdir -p gpurun_out
M=gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_final.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-canonical --no-e2e --no-sustained > /dev/null 2>&1
timeout 900 ncu --metrics $M --clock-control none --launch-skip 45 --launch-count 44 --csv --log-file gpurun_out/r02_step_ncu_metrics.csv python tools/ncu_targets.py fp32 3 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pool_fwd_kernel --launch-skip 1 --launch-count 1 -f -o gpurun_out/r02_pool_fwd python tools/ncu_targets.py fp32 2 > /dev/null 2>&1
ncu -i gpurun_out/r02_pool_fwd.ncu-rep --page raw --csv > gpurun_out/r02_pool_fwd_ncu_raw.csv 2>/dev/null
ls -la gpurun_out | tail -5
