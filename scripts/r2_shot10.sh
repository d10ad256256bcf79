#!/bin/bash
set -u
mkdir -p gpurun_out
run() { n=$1; shift; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 "$@"; }
F='CN=|MP_GPU|Error|error|assert'
{
  echo "== 2 GPUs nx 150 rep_min 64"; MP_NX=150 CTL_AMG_REP_MIN=64 run 2 tests/mp_gpu_check.py 2>&1 | grep -E "$F" | head
  echo "== 2 GPUs nx 150 default"; MP_NX=150 run 2 tests/mp_gpu_check.py 2>&1 | grep -E "$F" | head
} 2>&1 | tee gpurun_out/r2_mp2c.log
{
  run 2 scripts/mp_inner_time.py 2>&1 | grep inner_solve
  run 2 scripts/mp_inner_time.py 128 3 2>&1 | grep inner_solve
} 2>&1 | tee gpurun_out/r2_mp2_inner2.log
for wl in heat c3; do
  run 2 bench.py --gpus 2 --workload $wl --no_cpu_baseline --steps 2 --warmup 3 2> gpurun_out/r2_bench_${wl}_2gpu.err | grep '^{' > gpurun_out/r2_bench_${wl}_2gpu.json
  python -c "
import json; d=json.load(open('gpurun_out/r2_bench_${wl}_2gpu.json'))
print('$wl', {k:d.get(k) for k in ['value','iterations','kkt_residual','pc_apply_ms','kkt_apply_ms','setup_s']}, d['kernels']['inner_solve_ms'], d.get('alt_fgmres_triangular'))"
done
