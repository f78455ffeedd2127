#!/bin/bash
# round 2, GPU call V: pyramid / Scharr work items after the instruction diet (exact multiplies for /16 and /32, 32-bit item arithmetic)
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -x -q -m gpu -k "scharr or pyramid or lowpass or level or prepare or lucas or pyrlk or semi_dense or extruder or baseline" > gpurun_out/v_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/v_pytest.log
python tools/prep_bench.py
VPPB_PREPARE=streams python tools/prep_bench.py
timeout -k 10 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -s 20 -c 4 --csv --log-file gpurun_out/v_launches.csv python tools/prep_bench.py > /dev/null 2>&1
grep -E "k_pyrlk" gpurun_out/v_launches.csv | awk -F'","' '{print $5, $(NF-2), $NF}' | head -8
timeout -k 10 300 python bench.py --steps 3 --warmup 3 --cpu-budget 1 > gpurun_out/v_bench.json 2> gpurun_out/v_bench.err
python - <<PY
import json
d = json.loads(open('gpurun_out/v_bench.json').read().strip().splitlines()[-1])
print(json.dumps(d['extras']['pyrlk_1080p_10k'])[:300]); print(json.dumps(d['extras']['sdof_1080p'])[:200]); print(json.dumps(d['extras']['sdof_8k'])[:200])
PY
