#!/bin/bash
# round 2: N = 2 diagnosis of the tiled flow extra (9 ms where 2.3 ms were measured before): pyramid form and CPU binding varied
mkdir -p gpurun_out
for v in "A VPPB_PREPARE_UNSET=1 VPPB_BENCH_BIND=1" "B VPPB_PREPARE=streams VPPB_BENCH_BIND=1" "C VPPB_PREPARE_UNSET=1 VPPB_BENCH_BIND=0"; do
  set -- $v
  env $2 $3 timeout -k 10 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/n2c_$1.json 2> gpurun_out/n2c_$1.err
  python - <<PY
import json
d = json.loads(open('gpurun_out/n2c_$1.json').read().strip().splitlines()[-1])
x = d['extras']['sdof_8k_tiled']
print("$1 $2 $3", x.get('ms_per_frame_pair'), x.get('full_frame_agreement'), x.get('parity'), 'e2e', d['e2e']['value'], d['e2e'].get('cpus_bound_to_gpu_numa_node'))
PY
done
