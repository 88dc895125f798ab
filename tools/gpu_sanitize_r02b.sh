#!/bin/bash
# compute-sanitizer over what the second session of round 2 added: the register-tiled projector ("skinny") kernels, the launch path
# with programmatic dependent launch (every hot-path kernel starts with griddepcontrol.wait), the tile-fastest split-K wgrad order
mkdir -p gpurun_out
echo "=== memcheck"
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_kernels.py tests/test_gpu_gemm.py tests/test_gpu_model.py -q -m gpu \
  -k "(skinny or programmatic or tn_accum or tn_grouped or forward_train or cfg1) and not 1500" > gpurun_out/r02b_memcheck.log 2>&1
grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds|misaligned" gpurun_out/r02b_memcheck.log | sort | uniq -c | head -12
echo "=== racecheck (shared-memory staging of skinny_dgrad / skinny_wgrad)"
timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis python -m pytest tests/test_gpu_kernels.py -q -m gpu \
  -k "skinny and not 1500 and not 325" > gpurun_out/r02b_racecheck.log 2>&1
grep -E "RACECHECK SUMMARY|passed|failed|Race reported|hazard" gpurun_out/r02b_racecheck.log | sort | uniq -c | head -12
echo "=== synccheck"
timeout 600 compute-sanitizer --tool synccheck python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "skinny and not 1500" > gpurun_out/r02b_synccheck.log 2>&1
grep -E "ERROR SUMMARY|passed|failed|Barrier error|divergent" gpurun_out/r02b_synccheck.log | sort | uniq -c | head
