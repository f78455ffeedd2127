#!/bin/bash
# round 2: N = 8 line of bench.py with the final code (fused box tiles, row-tiled flow extra with the one-launch flow, agreement with the full frame)
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/n8_topo.txt 2>&1
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/n8_bench.json 2> gpurun_out/n8_bench.err; echo "bench n8 rc=$?"
tail -2 gpurun_out/n8_bench.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/n8_bench.json').read().strip().splitlines()[-1])
print({k: d.get(k) for k in ('value', 'ms_per_step', 'parity_checked', 'n_gpus')}, d['roofline']['frac'], d['e2e']['value'])
print(json.dumps(d.get('extras'))[:500])
PY
