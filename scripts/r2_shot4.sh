#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -p no:cacheprovider 2>&1 | tail -8 | tee gpurun_out/r2_gputests.log
{
  timeout 120 python scripts/inner_solve_time.py 2>&1 | tail -1
  CTL_SELL_FMT=dict8 timeout 120 python scripts/inner_solve_time.py 2>&1 | tail -1
  CTL_SELL_FMT_COARSE=f64 timeout 120 python scripts/inner_solve_time.py 2>&1 | tail -1
  CTL_SELL_FMT_COARSE=d16 timeout 120 python scripts/inner_solve_time.py 2>&1 | tail -1
  CTL_SELL_FMT_COARSE=pk timeout 120 python scripts/inner_solve_time.py 2>&1 | tail -1
  CTL_AMG_RR=0 timeout 120 python scripts/inner_solve_time.py 2>&1 | tail -1
  CTL_RR_LANES=8 timeout 120 python scripts/inner_solve_time.py 2>&1 | tail -1
  CTL_RR_LANES=16 timeout 120 python scripts/inner_solve_time.py 2>&1 | tail -1
  CTL_SELL_MIN_ROWS=10000 timeout 120 python scripts/inner_solve_time.py 2>&1 | tail -1
  CTL_SELL_MIN_ROWS=1000000 timeout 120 python scripts/inner_solve_time.py 2>&1 | tail -1
} | tee gpurun_out/r2_inner_variants2.log
timeout 600 ncu --cache-control none --clock-control none --metrics gpu__time_duration.sum --launch-skip 284 -c 142 --csv \
   --log-file gpurun_out/r2_inner_launches2.csv python scripts/inner_only.py 1024 8 > gpurun_out/r2_ncu1.log 2>&1
timeout 600 ncu --set full --import-source on --cache-control none --clock-control none -k regex:sell_cheb_kernel --launch-skip 24 -c 3 \
     -f -o gpurun_out/r2_cheb_stencil python scripts/inner_only.py 1024 3 > gpurun_out/r2_ncu_cheb_stencil.log 2>&1
timeout 600 python bench.py --no_cpu_baseline > gpurun_out/r2_bench_b.json 2> gpurun_out/r2_bench_b.err
tail -c 1500 gpurun_out/r2_bench_b.json
