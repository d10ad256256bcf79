#!/bin/bash
# round 2, call 21 (1 GPU): final verification -- the whole GPU suite, smoke(), inner solve, the bench lines of record
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests --maxfail=8 -q -m gpu -p no:cacheprovider 2>&1 | tail -40 | tee gpurun_out/r2_gputests21.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" 2>&1 | tail -4 | tee gpurun_out/r2_smoke21.log
timeout 200 python scripts/inner_solve_time.py 2>&1 | tail -1 | tee gpurun_out/r2_inner21.log
timeout 900 python bench.py 2> gpurun_out/r2_bench_c2_n1_final.err | grep '^{' > gpurun_out/r2_bench_c2_n1_final.json
timeout 900 python bench.py --workload c3 --no_cpu_baseline --steps 2 2> gpurun_out/r2_bench_c3_n1_final.err | grep '^{' > gpurun_out/r2_bench_c3_n1_final.json
python - <<'P'
import json
for f in ['r2_bench_c2_n1_final','r2_bench_c3_n1_final']:
    try:
        d=json.load(open(f'gpurun_out/{f}.json'))
        print(f, {k:d.get(k) for k in ['value','iterations','kkt_residual','pc_apply_ms','kkt_apply_ms','setup_s','gpu_launches','clocks']}, d['roofline']['frac'], d['roofline']['ms_per_launch'], d['roofline'].get('traffic'), d['roofline'].get('inner_solve'), d['roofline_spmm']['frac'], d.get('e2e'), d.get('alt_fgmres_triangular'), (d.get('cpu_baseline') or {}).get('value'), d.get('parity_vs_1gpu'))
    except Exception as e:
        print(f, 'ERR', e)
P
