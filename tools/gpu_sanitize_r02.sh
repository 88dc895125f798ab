#!/bin/bash
# compute-sanitizer over the kernels added or rewritten in round 2 (small shapes only: the tools slow kernels 10-100x):
# native executor paths, gated epilogue (pipelined TMEM drain, tiled gate stores), tiled gate_bwd, cp.async LayerNorm backward,
# InfoNCE over row lists, fp16 inference planes, Gram kernel, vectorised AdamW
mkdir -p gpurun_out
SEL='gated or gate_bwd or infonce_rows or f16 or ln_gelu or fused_adamw or smooth_rank or token_window or skip_missing or forward_train or n_views3 or cfg1'
echo "=== memcheck"
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_gemm.py tests/test_gpu_kernels.py tests/test_gpu_model.py tests/test_host_logic.py -q -m gpu \
  -k "($SEL) and not 1000 and not 30k" > gpurun_out/r02_memcheck.log 2>&1
grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds|misaligned" gpurun_out/r02_memcheck.log | sort | uniq -c | head -12
echo "=== racecheck (shared-memory staging of gate_bwd, cp.async stages and named barriers of ln_gelu_bwd, gated epilogue)"
timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis python -m pytest tests/test_gpu_gemm.py tests/test_gpu_kernels.py -q -m gpu \
  -k "(gate_bwd or dropout_forward_and_gate or ln_gelu_bwd or ln_gelu_dropout) and not 1000" > gpurun_out/r02_racecheck.log 2>&1
grep -E "RACECHECK SUMMARY|passed|failed|Race reported|hazard" gpurun_out/r02_racecheck.log | sort | uniq -c | head -12
grep -B2 -A12 "Race reported\|hazard" gpurun_out/r02_racecheck.log | grep -E "at .*mdl::|at .*at::" | sort | uniq -c | head -10
echo "=== synccheck"
timeout 600 compute-sanitizer --tool synccheck python -m pytest tests/test_gpu_gemm.py tests/test_gpu_kernels.py -q -m gpu \
  -k "(gate_bwd or dropout_forward_and_gate or ln_gelu or infonce_rows) and not 1000" > gpurun_out/r02_synccheck.log 2>&1
grep -E "ERROR SUMMARY|passed|failed|Barrier error|divergent" gpurun_out/r02_synccheck.log | sort | uniq -c | head
