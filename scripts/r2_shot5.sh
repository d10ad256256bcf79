#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -p no:cacheprovider 2>&1 | tail -8 | tee gpurun_out/r2_gputests.log
{
  timeout 120 python scripts/inner_solve_time.py 2>&1 | tail -1
  CTL_SELL_FMT_COARSE=stencil timeout 120 python scripts/inner_solve_time.py 2>&1 | tail -1
  CTL_SELL_FMT_COARSE=d16 timeout 120 python scripts/inner_solve_time.py 2>&1 | tail -1
  CTL_SELL_FMT=dict8 timeout 120 python scripts/inner_solve_time.py 2>&1 | tail -1
  CTL_SELL_FMT=f64 timeout 120 python scripts/inner_solve_time.py 2>&1 | tail -1
} | tee gpurun_out/r2_inner_variants3.log
timeout 600 ncu --cache-control none --clock-control none --metrics gpu__time_duration.sum --launch-skip 308 -c 154 --csv \
   --log-file gpurun_out/r2_inner_launches3.csv python scripts/inner_only.py 1024 8 > gpurun_out/r2_ncu1.log 2>&1
{
  for amg in "" "cycles=2,nu=4" "cycles=2,nu=5" "cycles=2,nu=6" "cycles=3,nu=2,nu_fine=3" "cycles=4,nu=2" "cycles=3,nu=3,nu_fine=4"; do
    echo "== fgmres triangular amg=[$amg]"
    timeout 300 python scripts/solve_c2.py --amg "$amg" 2>&1 | grep '"its"' | tail -1
    echo "== minres diagonal amg=[$amg]"
    timeout 300 python scripts/solve_c2.py --ksp minres --mode diagonal --amg "$amg" 2>&1 | grep '"its"' | tail -1
  done
} | tee gpurun_out/r2_amg_param_sweep.log
