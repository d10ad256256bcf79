#!/bin/bash
set -u
mkdir -p gpurun_out
run() { n=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 "$@"; }
F='CN=|MP_GPU|Error|error|assert'
{
  echo "== 1 GPU nx 150"; MP_NX=150 run 1 tests/mp_gpu_check.py 2>&1 | grep -E "$F" | head
  echo "== 2 GPUs default"; run 2 tests/mp_gpu_check.py 2>&1 | grep -E "$F" | head
  echo "== 2 GPUs rep_min 1"; CTL_AMG_REP_MIN=1 run 2 tests/mp_gpu_check.py 2>&1 | grep -E "$F" | head
  echo "== 2 GPUs nx 150 rep_min 64"; MP_NX=150 CTL_AMG_REP_MIN=64 run 2 tests/mp_gpu_check.py 2>&1 | grep -E "$F" | head
  echo "== 2 GPUs nx 150 rep_min 64 no graph"; CTL_NO_GRAPH=1 MP_NX=150 CTL_AMG_REP_MIN=64 run 2 tests/mp_gpu_check.py 2>&1 | grep -E "$F" | head
  echo "== 2 GPUs nx 150 default rep"; MP_NX=150 run 2 tests/mp_gpu_check.py 2>&1 | grep -E "$F" | head
  echo "== 2 GPUs nx 150 rep_min 16"; MP_NX=150 CTL_AMG_REP_MIN=16 run 2 tests/mp_gpu_check.py 2>&1 | grep -E "$F" | head
} 2>&1 | tee gpurun_out/r2_mp2b.log
{
  run 1 scripts/mp_inner_time.py 2>&1 | grep inner_solve
  run 2 scripts/mp_inner_time.py 2>&1 | grep inner_solve
  CTL_AMG_REP_MIN=100000000 run 2 scripts/mp_inner_time.py 2>&1 | grep inner_solve
  CTL_AMG_REP_MIN=4096 run 2 scripts/mp_inner_time.py 2>&1 | grep inner_solve
} 2>&1 | tee gpurun_out/r2_mp2_inner.log
run 2 bench.py --gpus 2 --no_cpu_baseline --steps 2 --warmup 3 2> gpurun_out/r2_bench_2gpu.err | grep '^{' > gpurun_out/r2_bench_2gpu.json
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_2gpu.json'))
print({k:d[k] for k in ['value','iterations','kkt_residual','pc_apply_ms','kkt_apply_ms','setup_s']}, d['kernels']['inner_solve_ms'], d['alt_fgmres_triangular'])"
