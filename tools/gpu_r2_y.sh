#!/bin/bash
# round 2, GPU call Y: flow SADs with the i1 window rows kept in registers across the batches of a match / an iteration
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -x -q -m gpu -k "semi_dense or sdof or extruder or flow or baseline" > gpurun_out/y_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/y_pytest.log
for occ in 2 3; do VPPB_SDOF_OCC=$occ timeout 300 python tools/sdof_bench.py 2>&1 | grep -v "schedule=\|^vppb_sdof" | tail -2; done
timeout -k 10 300 python bench.py --steps 3 --warmup 3 --cpu-budget 1 > gpurun_out/y_bench.json 2> gpurun_out/y_bench.err
python - <<PY
import json
d = json.loads(open('gpurun_out/y_bench.json').read().strip().splitlines()[-1])
print(json.dumps(d['extras']['sdof_1080p'])[:200]); print(json.dumps(d['extras']['sdof_8k'])[:200])
PY
