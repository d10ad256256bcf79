#!/bin/bash
# time one AMG inner solve (stream mode, warm) for different fused-tail grid sizes
for c in 8 16 32 64 148; do
  echo "CTL_FUSED_CTAS=$c"; CTL_FUSED_CTAS=$c python scripts/l2_persist.py 2>&1 | tail -1
done
echo "unfused"; CTL_NO_FUSED=1 python scripts/l2_persist.py 2>&1 | tail -1
