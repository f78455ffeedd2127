#!/bin/bash
# round 2, GPU call R: e2e leg with 4 / 6 / 8 frames in flight
mkdir -p gpurun_out
for n in 4 6 8 12; do
  VPPB_BENCH_INFLIGHT=$n timeout -k 10 300 python bench.py --steps 5 --warmup 3 --no-extras --cpu-budget 1 > gpurun_out/r_bench_$n.json 2> gpurun_out/r_bench_$n.err
  python - <<PY
import json
d = json.loads(open('gpurun_out/r_bench_$n.json').read().strip().splitlines()[-1])
print($n, d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['upload'], d['roofline']['frac'])
PY
done
