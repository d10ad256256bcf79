#!/bin/bash
# round 2, call 20 (1 GPU): the whole GPU suite with the new full-size tests, smoke(), CTA-size experiment
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests --maxfail=5 -q -m gpu -p no:cacheprovider --durations=8 2>&1 | tail -22 | tee gpurun_out/r2_gputests20.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" 2>&1 | tail -2 | tee gpurun_out/r2_smoke20.log
{
  timeout 200 python scripts/inner_solve_time.py 2>&1 | tail -1
  CTL_LIB_VARIANT=st256 timeout 200 python scripts/inner_solve_time.py 2>&1 | tail -1
} | tee gpurun_out/r2_inner20.log
