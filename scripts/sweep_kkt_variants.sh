#!/bin/bash
# Rebuild kkt_apply.cu with different (SCHUNK, SMINB) and time the kernel (run on the GPU box).
cd control_b200/csrc
for v in "4 3" "4 4" "8 2" "8 3"; do
  set -- $v
  rm -f ../lib/obj/kkt_apply.o
  make EXTRA="-DSCHUNK=$1 -DSMINB=$2" >/dev/null 2>&1
  echo "variant chunk=$1 minblocks=$2"
  (cd ../.. && python scripts/bench_apply.py --reps 20 2>&1 | tail -1)
done
rm -f ../lib/obj/kkt_apply.o; make >/dev/null 2>&1
cd ../.. && CTL_KKT_UNSTAGED=1 python scripts/bench_apply.py --reps 20 2>&1 | tail -1
