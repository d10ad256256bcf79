#!/bin/bash
# round 2, call 15: pre-wait matrix loads + speculative stencil gathers + device coarse inverse + new dense GEMV;
# parity first, then the inner solve, the set-up time, the AMG cycle parameters at C2 / C3, and a launch list
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests --maxfail=5 -q -m gpu -p no:cacheprovider 2>&1 | tail -15 | tee gpurun_out/r2_gputests16.log
{
  CTL_SETUP_TIMING=1 timeout 200 python scripts/inner_solve_time.py 2>gpurun_out/r2_setup_phases16.log | tail -1
  timeout 400 python scripts/mp_inner_time.py 128 3 2>&1 | grep inner_solve
} | tee gpurun_out/r2_inner16.log
{
  for amg in "" "cycles=2,nu=3" "cycles=2,nu=4" "cycles=1,nu=4" "cycles=1,nu=6" "cycles=2,nu=2"; do
    echo "== minres diagonal amg=[$amg]"
    timeout 300 python scripts/solve_c2.py --ksp minres --mode diagonal --amg "$amg" 2>&1 | grep '"its"' | tail -1
  done
  for amg in "" "cycles=2,nu=3"; do
    echo "== fgmres triangular amg=[$amg]"
    timeout 300 python scripts/solve_c2.py --amg "$amg" 2>&1 | grep '"its"' | tail -1
  done
  for amg in ""; do
    echo "== C3 gmres(10) triangular amg=[$amg]"
    timeout 400 python scripts/solve_c2.py --dim 3 --nx 128 --n_t 32 --be --ksp gmres --restart 10 --amg "$amg" 2>&1 | grep -E '"its"|pc setup' | tail -2
  done
} | tee gpurun_out/r2_amg_param_sweep16.log
timeout 600 ncu --cache-control none --clock-control none --metrics gpu__time_duration.sum --launch-skip 308 -c 154 --csv \
   --log-file gpurun_out/r2_inner_launches16.csv python scripts/inner_only.py 1024 8 > gpurun_out/r2_ncu16.log 2>&1
tail -2 gpurun_out/r2_ncu16.log
