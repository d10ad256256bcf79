#!/bin/bash
# round 2, call 18 (1 GPU): tests, the bench lines of record (C2 with the CPU leg, C3), format threshold experiment,
# launch list and ncu --set full captures of the final kernels
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests --maxfail=5 -q -m gpu -p no:cacheprovider 2>&1 | tail -6 | tee gpurun_out/r2_gputests18.log
timeout 900 python bench.py 2> gpurun_out/r2_bench_c2_n1.err | grep '^{' > gpurun_out/r2_bench_c2_n1.json
timeout 900 python bench.py --workload c3 --no_cpu_baseline --steps 2 2> gpurun_out/r2_bench_c3_n1.err | grep '^{' > gpurun_out/r2_bench_c3_n1.json
python - <<'P'
import json
for f in ['r2_bench_c2_n1','r2_bench_c3_n1']:
    try:
        d=json.load(open(f'gpurun_out/{f}.json'))
        print(f, {k:d.get(k) for k in ['value','iterations','kkt_residual','pc_apply_ms','kkt_apply_ms','setup_s','gpu_launches','clocks']}, d['roofline']['frac'], d['roofline']['ms_per_launch'], d['roofline'].get('inner_solve'), d['roofline_spmm']['frac'], d.get('e2e'), d.get('alt_fgmres_triangular'), (d.get('cpu_baseline') or {}).get('value'))
    except Exception as e:
        print(f, 'ERR', e)
P
{
  CTL_SETUP_TIMING=1 timeout 200 python scripts/inner_solve_time.py 2>gpurun_out/r2_setup_phases18.log | tail -1
  CTL_COARSE_PLAIN_MB=32 timeout 200 python scripts/inner_solve_time.py 2>&1 | tail -1
  CTL_COARSE_PLAIN_MB=16 timeout 200 python scripts/inner_solve_time.py 2>&1 | tail -1
} | tee gpurun_out/r2_inner18.log
grep "pc_setup" gpurun_out/r2_setup_phases18.log
timeout 600 ncu --cache-control none --clock-control none --metrics gpu__time_duration.sum -k regex:"sell_|csrv_|dense_gemv" --launch-skip 308 -c 154 --csv \
   --log-file gpurun_out/r2_inner_launches18.csv python scripts/inner_only.py 1024 8 > gpurun_out/r2_ncu18.log 2>&1
tail -1 gpurun_out/r2_ncu18.log
for k in sell_cheb_kernel dense_gemv_kernel sell_spmv_kernel; do
  timeout 300 ncu --set full --cache-control none --clock-control none --import-source on -k regex:$k --launch-skip 100 -c 3 \
     -o gpurun_out/r2_full18_$k -f python scripts/inner_only.py 1024 6 > gpurun_out/r2_ncu_full18_$k.log 2>&1
done
# the fine-level smoother with COLD caches (ncu flushes before every replay): DRAM traffic per launch
timeout 300 ncu --set full --cache-control all --clock-control none -k regex:sell_cheb_kernel --launch-skip 100 -c 3 \
     -o gpurun_out/r2_full18_sell_cheb_cold -f python scripts/inner_only.py 1024 6 > gpurun_out/r2_ncu_full18_cold.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -5
