#!/bin/bash
# per-kernel durations of one inner solve with warm caches, and full ncu sections of the four kernel types
set -u
mkdir -p gpurun_out
timeout 300 python scripts/inner_only.py 1024 8 > gpurun_out/r2_inner_only.log 2>&1 || { tail -5 gpurun_out/r2_inner_only.log; exit 1; }
cat gpurun_out/r2_inner_only.log | tail -2
timeout 600 ncu --cache-control none --clock-control none --metrics gpu__time_duration.sum --launch-skip 284 -c 142 --csv \
   --log-file gpurun_out/r2_inner_launches.csv python scripts/inner_only.py 1024 8 > gpurun_out/r2_ncu1.log 2>&1
for k in sell_cheb_kernel sell_first2_kernel csrv_rr_kernel dense_gemv_kernel csrv_cheb_kernel sell_spmv_kernel csrv_spmv_kernel; do
  timeout 600 ncu --set full --import-source on --cache-control none --clock-control none -k regex:$k --launch-skip 12 -c 1 \
     -f -o gpurun_out/r2_$k python scripts/inner_only.py 1024 3 > gpurun_out/r2_ncu_$k.log 2>&1
done
ls -la gpurun_out/*.ncu-rep | tail
