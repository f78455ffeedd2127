#!/bin/bash
# round 2, GPU call W: semi-dense flow sweeps that skip the SAD batch when no neighbour can become a candidate; N4 / flow tests; ncu of the fused launch at 8K
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -x -q -m gpu -k "semi_dense or sdof or extruder or flow or baseline or n4 or oriented or prepare" > gpurun_out/w_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/w_pytest.log
VPPB_SDOF_STATS=1 timeout 300 python tools/sdof_bench.py 2>&1 | grep -v "schedule=" | sort | uniq -c | tail -6
timeout -k 10 300 python bench.py --steps 3 --warmup 3 --cpu-budget 1 > gpurun_out/w_bench.json 2> gpurun_out/w_bench.err
python - <<PY
import json
d = json.loads(open('gpurun_out/w_bench.json').read().strip().splitlines()[-1])
print(json.dumps(d['extras']['sdof_1080p'])[:200]); print(json.dumps(d['extras']['sdof_8k'])[:200]); print(json.dumps(d['extras']['pyrlk_1080p_10k'])[:200])
PY
sed -n '/^cat > \/tmp\/sdof1.py/,/^PY$/p' tools/gpu_r2_m.sh | sed '1d;$d' > /tmp/sdof1.py
timeout -k 10 400 ncu --set full --clock-control none --import-source on -k regex:k_sdof_fused -s 1 -c 1 -f -o gpurun_out/w_prof_sdof python /tmp/sdof1.py > gpurun_out/w_ncu_sdof.log 2>&1
ls -la gpurun_out/w_prof_sdof.ncu-rep
