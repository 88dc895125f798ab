#!/bin/bash
# 2-GPU validation: torchrun bench (weak scaling) for both arms, as the driver launches it
mkdir -p gpurun_out
nvidia-smi -L
echo "=== bench 2 gpus"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 2>&1 | tee gpurun_out/bench_2gpu.log | tail -2 | cut -c1-1500
echo "=== reference arm 2 gpus"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>&1 | tail -1 | cut -c1-400
echo "=== bench 1 gpu (same box)"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-600
