#!/bin/bash
# usage: r2_shot19.sh N [workloads...] -- parity of the distributed sweeps on N ranks, then the named bench workloads
set -u
N=$1; shift
mkdir -p gpurun_out
run() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 "$@"; }
MP_NX=150 run tests/mp_gpu_check.py 2>&1 | grep -E "CN=|MP_GPU|Error|error|assert" | head | tee gpurun_out/r2_parity_n$N.log
for wl in "$@"; do
  run bench.py --gpus $N --workload $wl --no_cpu_baseline --steps 2 --warmup 3 2> gpurun_out/r2_bench_${wl}_n$N.err | grep '^{' > gpurun_out/r2_bench_${wl}_n$N.json
  python -c "
import json; d=json.load(open('gpurun_out/r2_bench_${wl}_n$N.json'))
print('$wl', 'N=$N', {k:d.get(k) for k in ['value','iterations','kkt_residual','pc_apply_ms','kkt_apply_ms','setup_s','clocks','parity_vs_1gpu']}, d['kernels']['inner_solve_ms'], d.get('alt_fgmres_triangular'))" || tail -5 gpurun_out/r2_bench_${wl}_n$N.err
done
