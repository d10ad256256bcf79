#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -p no:cacheprovider 2>&1 | tail -6 | tee gpurun_out/r2_gputests.log
{
timeout 400 python scripts/mp_inner_time.py 128 3 2>&1 | grep inner_solve
timeout 400 python scripts/mp_inner_time.py 96 3 2>&1 | grep inner_solve
} | tee gpurun_out/r2_inner_c3b.log
timeout 600 ncu --cache-control none --clock-control none --metrics gpu__time_duration.sum --launch-skip 400 -c 80 --csv \
   --log-file gpurun_out/r2_inner_launches_c3b.csv python scripts/mp_inner_time.py 128 3 > gpurun_out/r2_ncu_c3.log 2>&1
