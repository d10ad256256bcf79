#!/bin/bash
# usage: r2_scale.sh N  -- parity of the distributed sweeps on N ranks, then config C3 and config C2 on N GPUs
set -u
N=$1
mkdir -p gpurun_out
run() { timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 "$@"; }
MP_NX=150 run tests/mp_gpu_check.py 2>&1 | grep -E "CN=|MP_GPU|Error|error|assert" | head | tee gpurun_out/r2_scale_parity_n$N.log
for wl in c3 heat; do
  run bench.py --gpus $N --workload $wl --no_cpu_baseline --steps 2 --warmup 3 2> gpurun_out/r2_scale_${wl}_n$N.err | grep '^{' > gpurun_out/r2_scale_${wl}_n$N.json
  python -c "
import json; d=json.load(open('gpurun_out/r2_scale_${wl}_n$N.json'))
print('$wl', 'N=$N', {k:d.get(k) for k in ['value','iterations','kkt_residual','pc_apply_ms','kkt_apply_ms','setup_s','clocks']}, d['kernels']['inner_solve_ms'], d.get('alt_fgmres_triangular'))" || tail -5 gpurun_out/r2_scale_${wl}_n$N.err
done
