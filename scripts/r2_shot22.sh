#!/bin/bash
# round 2, call 22 (2 GPUs): distributed parity incl. ctl_build_rhs / ctl_objective on two ranks
set -u
N=2
mkdir -p gpurun_out
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 "$@"; }
{ MP_NX=60 run tests/mp_gpu_check.py 2>&1 | grep -E "CN=|MP_GPU|Error|error|assert|Traceback" | head -20; } | tee gpurun_out/r2_mp22.log
