#!/bin/bash
# round 2, last GPU call: LK matching kernel with 4 / 2 / 1 warps per CTA (SM balance of a one-wave, issue-bound launch)
mkdir -p gpurun_out
timeout 200 python tools/lk_bench.py 2>&1 | tail -8 | tee gpurun_out/lk_bench.txt
