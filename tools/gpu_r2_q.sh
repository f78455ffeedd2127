#!/bin/bash
# round 2, GPU call Q: N4 kernels after tuning (lbp row prefetch, local_maxima_filter runs), host link ceiling, bench extras
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_n4.py tests/test_cpp_api.py -x -q -m gpu > gpurun_out/q_pytest_n4.log 2>&1; echo "pytest n4 rc=$?"; tail -3 gpurun_out/q_pytest_n4.log
timeout 300 python tools/pcie_probe.py > gpurun_out/q_pcie.json 2> gpurun_out/q_pcie.err; cat gpurun_out/q_pcie.json
timeout -k 10 600 python bench.py --steps 5 --warmup 3 --cpu-budget 2 > gpurun_out/q_bench_n1.json 2> gpurun_out/q_bench_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/q_bench_n1.json').read().strip().splitlines()[-1])
print({k: d.get(k) for k in ('value', 'ms_per_step')}, d['roofline']['frac'], d['roofline'].get('traffic'), json.dumps(d['e2e'])[:400])
for k in ('lbp_u8_4k', 'local_maxima_filter_1080p'):
    print(k, json.dumps(d['extras'][k])[:330])
PY
