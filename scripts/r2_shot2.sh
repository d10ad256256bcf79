#!/bin/bash
# Round 2, second GPU call: GPU parity tests with the rewritten sweep kernels, then inner-solve A/B timings.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -p no:cacheprovider 2>&1 | tail -25 | tee gpurun_out/r2_gputests.log
{
  for fmt in f64 d16 pk dict16 dict8; do
    CTL_SELL_FMT=$fmt timeout 120 python scripts/inner_solve_time.py 2>&1 | tail -1
  done
  CTL_AMG_RR=0 timeout 120 python scripts/inner_solve_time.py 2>&1 | tail -1
  CTL_NO_PDL=1 timeout 120 python scripts/inner_solve_time.py 2>&1 | tail -1
  timeout 120 python scripts/inner_solve_time.py 1024 '{"coarse_max": 600}' 2>&1 | tail -1
  CTL_SELL_MIN_ROWS=10000 timeout 120 python scripts/inner_solve_time.py 2>&1 | tail -1
  CTL_CSR_LANES_SHIFT=1 timeout 120 python scripts/inner_solve_time.py 2>&1 | tail -1
} | tee gpurun_out/r2_inner_variants.log
timeout 600 python bench.py --no_cpu_baseline > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err
tail -c 3000 gpurun_out/r2_bench_a.json
