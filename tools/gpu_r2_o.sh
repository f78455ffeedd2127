#!/bin/bash
# round 2, GPU call O: semi-dense flow as ONE cooperative launch (relaxation schedule): parity, timing, statistics, launch list
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -m gpu -x -q -k "sdof or semi_dense or extruder or flow or baseline" > gpurun_out/o_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/o_pytest.log
timeout -k 10 600 python bench.py --steps 5 --warmup 3 --cpu-budget 2 > gpurun_out/o_bench_n1.json 2> gpurun_out/o_bench_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/o_bench_n1.json').read().strip().splitlines()[-1])
for k in ('sdof_1080p', 'sdof_8k'):
    print(k, json.dumps(d['extras'][k])[:400])
print('cpu sdof', json.dumps(d['extras']['cpu'].get('sdof_1080p')))
PY
sed -n '/^cat > \/tmp\/sdof1.py/,/^PY$/p' tools/gpu_r2_m.sh | sed '1d;$d' > /tmp/sdof1.py
VPPB_SDOF_STATS=1 timeout 120 python /tmp/sdof1.py 2>&1 | tail -4
timeout -k 10 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -c 200 --csv --log-file gpurun_out/o_sdof_launches.csv python /tmp/sdof1.py > /dev/null 2>&1
grep k_sdof gpurun_out/o_sdof_launches.csv | awk -F'","' '{print $5, $(NF-2), $NF}' | tail -12
VPPB_SDOF_STATS=1 timeout 300 python tools/sdof_bench.py 2>&1 | grep -v "^vppb_sdof_u8" | tail -8
