#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -p no:cacheprovider 2>&1 | tail -6 | tee gpurun_out/r2_gputests.log
timeout 120 python scripts/inner_solve_time.py 2>&1 | tail -1 | tee gpurun_out/r2_inner_now.log
timeout 900 python bench.py --workload c3 --steps 2 --warmup 3 2> gpurun_out/r2_bench_c3_1gpu.err | grep '^{' > gpurun_out/r2_bench_c3_1gpu.json
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_c3_1gpu.json'))
print({k:d.get(k) for k in ['value','iterations','kkt_residual','pc_apply_ms','kkt_apply_ms','setup_s','gpu_launches']}, d.get('kernels'), d.get('roofline_spmm',{}).get('frac'))"
tail -3 gpurun_out/r2_bench_c3_1gpu.err
