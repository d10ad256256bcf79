#!/bin/bash
set -u
mkdir -p gpurun_out
{
timeout 200 python scripts/inner_solve_time.py 2>&1 | tail -1
timeout 400 python scripts/mp_inner_time.py 128 3 2>&1 | grep inner_solve
CTL_COARSE_PLAIN_MB=100000 timeout 400 python scripts/mp_inner_time.py 128 3 2>&1 | grep inner_solve
CTL_SELL_FMT=dict16 timeout 400 python scripts/mp_inner_time.py 128 3 2>&1 | grep inner_solve
} | tee gpurun_out/r2_inner_c3.log
timeout 900 python bench.py --workload c3 --steps 2 --warmup 3 2> gpurun_out/r2_bench_c3_1gpu.err | grep '^{' > gpurun_out/r2_bench_c3_1gpu.json
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_c3_1gpu.json'))
print({k:d.get(k) for k in ['value','iterations','kkt_residual','pc_apply_ms','kkt_apply_ms','setup_s','gpu_launches']}, d.get('kernels'))"
timeout 600 ncu --cache-control none --clock-control none --metrics gpu__time_duration.sum --launch-skip 400 -c 160 --csv \
   --log-file gpurun_out/r2_inner_launches_c3.csv python scripts/mp_inner_time.py 128 3 > gpurun_out/r2_ncu_c3.log 2>&1
