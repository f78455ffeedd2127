#!/bin/bash
# round 2, GPU call K: parity, bench N=1, FAST launch split + ncu of band and emit kernels
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests -x -q -m gpu > gpurun_out/k_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/k_pytest.log
timeout -k 10 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/k_bench_n1.json 2> gpurun_out/k_bench_n1.err
timeout -k 10 600 python tools/fast_bench.py > gpurun_out/k_fast_bench.json 2> gpurun_out/k_fast_bench.err
timeout -k 10 900 python tools/box_bench.py > gpurun_out/k_box_bench.json 2> gpurun_out/k_box_bench.err
cat > /tmp/fast4k.py <<'PY'
import sys, numpy as np
sys.path.insert(0, '.')
import vpp_b200 as vpp
from vpp_b200 import capi
from tests import scenes
capi.check(capi.lib.vppb_init(0))
G = vpp.Image2d.from_host(scenes.rectangles_scene(2160, 3840, seed=42), "u8", border=3); vpp.fill_border_mirror(G)
for _ in range(3): k = vpp.fast9(G, 20)
print(len(k))
PY
timeout -k 10 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/k_fast_launches.csv python /tmp/fast4k.py > /dev/null 2>&1
timeout -k 10 300 ncu --set full --clock-control none --import-source on -k regex:k_fast9 -s 2 -c 2 -f -o gpurun_out/k_prof_fast4k python /tmp/fast4k.py > gpurun_out/k_ncu_fast.log 2>&1
tail -3 gpurun_out/k_pytest.log
python - <<'PY'
import json
try:
    l = json.loads(open('gpurun_out/k_bench_n1.json').read().strip().splitlines()[-1])
    print({k: l[k] for k in ('value', 'ms_per_step', 'parity_checked', 'gpu_launches')}, l['roofline']['frac'], l['e2e']['value'], l['clocks'])
    for k, v in l.get('extras', {}).items():
        print(k, json.dumps(v)[:300])
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/k_bench_n1.err').read()[-2000:])
PY
tail -7 gpurun_out/k_box_bench.err
tail -3 gpurun_out/k_fast_bench.err
grep k_fast9 gpurun_out/k_fast_launches.csv | tail -4 | awk -F'","' '{print $5, $NF}'
