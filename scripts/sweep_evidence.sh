#!/bin/bash
# BASELINE config #5 lines (DDIM steps x clips per step) + a clean GPU test log; run through gpurun from the repo root.
set -u
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > $O/r02g_gpu_tests.log 2>&1; tail -2 $O/r02g_gpu_tests.log
: > $O/r02g_bench_sweep.log
for cfg in "--ddim-steps 10" "--ddim-steps 25" "--ddim-steps 100 --steps 2" "--ddim-steps 25 --clips-per-step 4 --steps 2"; do
  echo "# python bench.py $cfg --warmup 3 --no-library-baseline --no-cpu-baseline" >> $O/r02g_bench_sweep.log
  timeout 900 python bench.py $cfg --warmup 3 --no-library-baseline --no-cpu-baseline 2>/dev/null | tail -1 >> $O/r02g_bench_sweep.log
  tail -1 $O/r02g_bench_sweep.log | cut -c1-200
done
