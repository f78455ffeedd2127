#!/bin/bash
# round 2, GPU call E: parity, the new bench (N=1, with extras) + reference arm, ncu traffic capture of the bench regime, FAST timing
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests -x -q -m gpu > gpurun_out/e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/e_pytest.log
timeout -k 10 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/e_bench_n1.json 2> gpurun_out/e_bench_n1.err
timeout -k 10 300 python bench.py --impl reference --gpus 1 --steps 5 --warmup 2 > gpurun_out/e_bench_ref.json 2> gpurun_out/e_bench_ref.err
timeout -k 10 600 python tools/fast_bench.py > gpurun_out/e_fast_bench.json 2> gpurun_out/e_fast_bench.err
# DRAM traffic of the streaming kernel in the bench regime: steady state (skip 40 launches), caches left as the bench leaves them
timeout -k 10 400 ncu --set full --cache-control none --clock-control none --import-source on -k regex:k_box5_stream -s 40 -c 2 -f -o gpurun_out/e_prof_box_bench python bench.py --gpus 1 --steps 2 --warmup 3 --no-extras --graph 0 --passes 16 > gpurun_out/e_ncu_box.log 2>&1
timeout -k 10 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 200 --csv --log-file gpurun_out/e_bench_launches.csv python bench.py --gpus 1 --steps 2 --warmup 3 --no-extras --graph 0 --passes 8 > /dev/null 2>&1
tail -3 gpurun_out/e_pytest.log
python - <<'PY'
import json
try:
    l = json.loads(open('gpurun_out/e_bench_n1.json').read().strip().splitlines()[-1])
    print({k: l[k] for k in ('value', 'ms_per_step', 'parity_checked', 'gpu_launches')}, l['roofline']['frac'], l['e2e']['value'], l['clocks'])
    for k, v in l.get('extras', {}).items():
        print(k, json.dumps(v)[:400])
    print(l.get('cpu_baseline'))
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/e_bench_n1.err').read()[-2000:])
PY
tail -3 gpurun_out/e_fast_bench.err
