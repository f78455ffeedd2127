#!/bin/bash
# round 2, GPU call U: LK v2 kernel after the instruction diet (register-resident sampler, integer-pipe byte->float, 128-bit ordered sums)
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -x -q -m gpu -k "lucas or pyrlk or lk or baseline or cpp or prepare or semi_dense" > gpurun_out/u_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/u_pytest.log
timeout -k 10 300 python bench.py --steps 3 --warmup 3 --cpu-budget 1 > gpurun_out/u_bench.json 2> gpurun_out/u_bench.err
python - <<PY
import json
d = json.loads(open('gpurun_out/u_bench.json').read().strip().splitlines()[-1])
print(json.dumps(d['extras']['pyrlk_1080p_10k'])[:300]); print(json.dumps(d['extras']['sdof_1080p'])[:200]); print(json.dumps(d['extras']['cpu'].get('pyrlk_1080p_10k')))
PY
