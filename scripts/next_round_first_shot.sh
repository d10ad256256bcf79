#!/bin/bash
# First GPU call of the next round (one B200, about 2 minutes): the KKT-apply experiments that were written at the
# end of round 1 without GPU time left.  Usage:
#   gpurun --timeout 300 -- 'bash scripts/next_round_first_shot.sh'
# Output: gpurun_out/kkt_variants.log (bit-exactness against the default kernel + timings at 512^2 and 1024^2) and
# gpurun_out/kkt_variants_tests.log (the opt-in multi-shape tests).
set -u
mkdir -p gpurun_out
{
  for nx in 512 1024; do
    for env in CTL_KKT_TMA=4 "CTL_KKT_TMA=4 --env CTL_TILE_ROWS=16" CTL_KKT_TMA=3 "CTL_KKT_TMA=3 --env CTL_TILE_ROWS=16" CTL_KKT_TMA=2 "CTL_KKT_TMA=2 --env CTL_TILE_ROWS=16"; do
      timeout 60 python scripts/compare_apply_variants.py --nx $nx --env $env 2>&1 | tail -1
    done
  done
  timeout 60 python scripts/compare_apply_variants.py --nx 1024 --be --env CTL_KKT_TMA=4 2>&1 | tail -1
} | tee gpurun_out/kkt_variants.log
CTL_RUN_UNVERIFIED=1 timeout 180 python -m pytest tests/test_zz_gpu_late.py -v -m gpu --no-header -p no:cacheprovider -k pipelined 2>&1 \
  | tee gpurun_out/kkt_variants_tests.log | tail -15
