#!/bin/bash
set -u
mkdir -p gpurun_out
{
timeout 400 python scripts/mp_inner_time.py 128 3 2>&1 | grep inner_solve
timeout 200 python scripts/inner_solve_time.py 2>&1 | tail -1
} | tee gpurun_out/r2_inner_c3c.log
nproc; lscpu | grep -E "Model name|Socket|Core|Thread" 
OMP_WAIT_POLICY=active timeout 900 python bench.py --impl reference --steps 1 --warmup 1 2> gpurun_out/r2_ref_c2.err | grep '^{' > gpurun_out/r2_ref_c2.json
python -c "
import json; d=json.load(open('gpurun_out/r2_ref_c2.json')); print(d['value'], d['cpu_baseline'])"
tail -3 gpurun_out/r2_ref_c2.err
