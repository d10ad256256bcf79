#!/bin/bash
# round 2, call 17 (2 GPUs): parity of the distributed sweeps with the pre-wait loads on the multi-GPU path, inner
# solve with and without them, config C2 on 2 GPUs
set -u
N=2
mkdir -p gpurun_out
run() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 "$@"; }
{
MP_NX=150 run tests/mp_gpu_check.py 2>&1 | grep -E "CN=|MP_GPU|Error|error|assert" | head
run scripts/mp_inner_time.py 1024 2 2>&1 | grep inner_solve
CTL_MP_EARLY_WAIT=1 run scripts/mp_inner_time.py 1024 2 2>&1 | grep inner_solve
} | tee gpurun_out/r2_mp17.log
run bench.py --gpus $N --no_cpu_baseline --steps 2 --warmup 3 2> gpurun_out/r2_bench_heat_n2.err | grep '^{' > gpurun_out/r2_bench_heat_n2.json
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_heat_n2.json'))
print({k:d.get(k) for k in ['value','iterations','kkt_residual','pc_apply_ms','kkt_apply_ms','setup_s','clocks','parity_vs_1gpu']}, d['kernels']['inner_solve_ms'], d.get('alt_fgmres_triangular'))" || tail -5 gpurun_out/r2_bench_heat_n2.err
