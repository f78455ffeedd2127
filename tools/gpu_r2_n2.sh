#!/bin/bash
# round 2: N = 2 check of bench.py after the semi-dense flow became one cooperative launch (row-tiled flow extra, fused box tiles, tile tests)
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/n2_topo.txt 2>&1; head -14 gpurun_out/n2_topo.txt
lscpu | grep -E 'Model name|Socket|NUMA|^CPU\(s\)' > gpurun_out/n2_lscpu.txt 2>&1; cat gpurun_out/n2_lscpu.txt
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/n2_bench.json 2> gpurun_out/n2_bench.err; echo "bench n2 rc=$?"
tail -3 gpurun_out/n2_bench.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/n2_bench.json').read().strip().splitlines()[-1])
print({k: d.get(k) for k in ('value', 'ms_per_step', 'parity_checked', 'n_gpus')}, d['roofline']['frac'], d['e2e']['value'])
print(json.dumps(d.get('extras'))[:600])
PY
