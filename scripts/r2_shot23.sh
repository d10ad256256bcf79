#!/bin/bash
# round 2, call 23 (1 GPU): the whole GPU suite once more (multi-rank rhs / objective code paths on one rank, Stokes P=
# hook, full-size C4 / C5 tests)
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests --maxfail=8 -q -m gpu -p no:cacheprovider --durations=12 2>&1 | tail -45 | tee gpurun_out/r2_gputests23.log
