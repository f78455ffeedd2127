#!/bin/bash
# round 2, GPU call P: N4 (lbp, local_maxima_filter, blockwise_rank, oriented LK) on hardware, the whole GPU suite, the fused flow kernel at 2 / 3 / 4 CTAs per SM, bench
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_n4.py -x -q -m gpu > gpurun_out/p_pytest_n4.log 2>&1; echo "pytest n4 rc=$?"; tail -3 gpurun_out/p_pytest_n4.log
timeout -k 10 1200 python -m pytest tests -x -q -m gpu > gpurun_out/p_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/p_pytest.log
for occ in 2 3 4; do echo "VPPB_SDOF_OCC=$occ"; VPPB_SDOF_OCC=$occ VPPB_SDOF_STATS=1 timeout 300 python tools/sdof_bench.py 2>&1 | grep -v "schedule=" | sort | uniq -c | tail -6; done
timeout -k 10 600 python bench.py --steps 5 --warmup 3 --cpu-budget 2 > gpurun_out/p_bench_n1.json 2> gpurun_out/p_bench_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/p_bench_n1.json').read().strip().splitlines()[-1])
print({k: d.get(k) for k in ('value', 'ms_per_step')}, d['roofline']['frac'], d['e2e']['value'])
for k, v in d['extras'].items():
    print(k, json.dumps(v)[:330])
PY
