#!/bin/bash
# round 2, GPU call X: flow matching with the prediction's SAD in the first descent batch; whole GPU suite
mkdir -p gpurun_out
timeout -k 10 1200 python -m pytest tests -x -q -m gpu > gpurun_out/x_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/x_pytest.log
VPPB_SDOF_STATS=1 timeout 300 python tools/sdof_bench.py 2>&1 | grep -v "schedule=\|^vppb_sdof" | tail -3
timeout -k 10 300 python bench.py --steps 3 --warmup 3 --cpu-budget 1 > gpurun_out/x_bench.json 2> gpurun_out/x_bench.err
python - <<PY
import json
d = json.loads(open('gpurun_out/x_bench.json').read().strip().splitlines()[-1])
print(json.dumps(d['extras']['sdof_1080p'])[:200]); print(json.dumps(d['extras']['sdof_8k'])[:200])
PY
