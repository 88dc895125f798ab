#!/bin/bash
# compute-sanitizer over the kernels added or rewritten at the end of round 1 (small shapes only: the tools slow kernels 10-100x)
mkdir -p gpurun_out
SEL='ln_gelu or gather_rows or bag_sums or row_indexed'
echo "=== memcheck (LayerNorm kernels, row gather, sampler, big GOT)"
timeout 400 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_kernels.py tests/test_datasets.py tests/test_gpu_got.py -q -m gpu \
  -k "($SEL or sample_gather or match_shared_memory or 2-97) and not 1000" > gpurun_out/memcheck2.log 2>&1
grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds|misaligned" gpurun_out/memcheck2.log | sort | uniq -c | head -12
echo "=== racecheck (named-barrier LayerNorm backward, big GOT staging buffers)"
timeout 500 compute-sanitizer --tool racecheck --racecheck-report analysis python -m pytest tests/test_gpu_kernels.py tests/test_gpu_got.py -q -m gpu \
  -k "(($SEL) and not 1000 and not fwd) or 3-6" > gpurun_out/racecheck2.log 2>&1
grep -E "RACECHECK SUMMARY|passed|failed|Race reported|hazard" gpurun_out/racecheck2.log | sort | uniq -c | head -12
grep -B2 -A12 "Race reported\|hazard" gpurun_out/racecheck2.log | grep -E "at .*mdl::|at .*at::" | sort | uniq -c | head -10
echo "=== synccheck"
timeout 300 compute-sanitizer --tool synccheck python -m pytest tests/test_gpu_kernels.py tests/test_gpu_got.py -q -m gpu \
  -k "(($SEL) and not 1000) or 3-6 or 2-97" > gpurun_out/synccheck2.log 2>&1
grep -E "ERROR SUMMARY|passed|failed|Barrier error|divergent" gpurun_out/synccheck2.log | sort | uniq -c | head
