#!/bin/bash
# round 2, GPU call S: vppb_pyrlk_prepare as one cooperative launch (parity with the per-step launches, LK suite, timing), whole suite, bench
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -x -q -m gpu > gpurun_out/s_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/s_pytest.log
for form in streams fused; do
  VPPB_PREPARE=$form timeout -k 10 300 python bench.py --steps 3 --warmup 3 --cpu-budget 1 > gpurun_out/s_bench_$form.json 2> gpurun_out/s_bench_$form.err
  python - <<PY
import json
d = json.loads(open('gpurun_out/s_bench_$form.json').read().strip().splitlines()[-1])
print("$form", json.dumps(d['extras']['pyrlk_1080p_10k'])[:300], d['extras']['sdof_1080p']['ms_pyramids'], d['e2e']['value'])
PY
done
