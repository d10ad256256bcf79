#!/bin/bash
# 2 GPUs: parity of the distributed sweeps against the single-rank oracle, then timings
set -u
mkdir -p gpurun_out
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 "$@"; }
{
  echo "== default rep_min"; run tests/mp_gpu_check.py 2>&1 | grep -E "CN=|MP_GPU|Error|error|assert" | head -20
  echo "== rep_min 1 (every level but the dense one distributed)"; CTL_AMG_REP_MIN=1 run tests/mp_gpu_check.py 2>&1 | grep -E "CN=|MP_GPU|Error|error|assert" | head -20
  echo "== nx 96, rep_min 64"; MP_NX=96 CTL_AMG_REP_MIN=64 run tests/mp_gpu_check.py 2>&1 | grep -E "CN=|MP_GPU|Error|error|assert" | head -20
  echo "== no graph"; CTL_NO_GRAPH=1 CTL_AMG_REP_MIN=1 run tests/mp_gpu_check.py 2>&1 | grep -E "CN=|MP_GPU|Error|error|assert" | head -20
} 2>&1 | tee gpurun_out/r2_mp2.log
run bench.py --gpus 2 --no_cpu_baseline --steps 2 --warmup 3 > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/r2_bench_2gpu.err
tail -c 2500 gpurun_out/r2_bench_2gpu.json; tail -5 gpurun_out/r2_bench_2gpu.err
