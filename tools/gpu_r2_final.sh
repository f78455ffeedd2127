#!/bin/bash
# round 2, final single-GPU call: everything the driver runs (tests, smoke, both bench arms) + the ncu captures the profiles/ summaries are made from
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > gpurun_out/z_smi.txt 2>&1
timeout -k 10 600 python -m pytest tests -x -q -m gpu > gpurun_out/z_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/z_pytest.log
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/z_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/z_smoke.log
timeout -k 10 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/z_bench_ref.json 2> gpurun_out/z_bench_ref.err
timeout -k 10 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/z_bench_n1.json 2> gpurun_out/z_bench_n1.err
# ncu: DRAM traffic + issue statistics of the streaming kernel in the bench regime (128-frame launches, steady state, caches as the bench leaves them)
timeout -k 10 400 ncu --set full --cache-control none --clock-control none --import-source on -k regex:k_box5_stream -s 12 -c 2 -f -o gpurun_out/z_prof_box_bench python bench.py --gpus 1 --steps 2 --warmup 3 --no-extras --graph 0 --passes 4 > gpurun_out/z_ncu_box.log 2>&1
# launch list of the same command: the kernel's share of the step
timeout -k 10 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 8 -c 60 --csv --log-file gpurun_out/z_bench_launches.csv python bench.py --gpus 1 --steps 2 --warmup 3 --no-extras --graph 0 --passes 4 > /dev/null 2>&1
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
timeout -k 10 300 ncu --set full --clock-control none --import-source on -k regex:k_fast9 -s 2 -c 2 -f -o gpurun_out/z_prof_fast4k python /tmp/fast4k.py > gpurun_out/z_ncu_fast.log 2>&1
timeout -k 10 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/z_fast_launches.csv python /tmp/fast4k.py > /dev/null 2>&1
# semi-dense flow: ncu of the single cooperative launch (1080p, a keypoint in every 10 x 10 block), its relaxation statistics, all schedules timed
sed -n '/^cat > \/tmp\/sdof1.py/,/^PY$/p' tools/gpu_r2_m.sh | sed '1d;$d' > /tmp/sdof1.py
timeout -k 10 400 ncu --set full --clock-control none --import-source on -k regex:k_sdof_fused -s 1 -c 1 -f -o gpurun_out/z_prof_sdof python /tmp/sdof1.py > gpurun_out/z_ncu_sdof.log 2>&1
VPPB_SDOF_STATS=1 timeout 120 python /tmp/sdof1.py > /dev/null 2> gpurun_out/z_sdof_stats.txt
timeout 300 python tools/sdof_bench.py > gpurun_out/z_sdof_bench.txt 2>&1
timeout 300 python tools/pcie_probe.py > gpurun_out/z_pcie.json 2> gpurun_out/z_pcie.err
timeout 120 tests/cpp/_build/pw_bench > gpurun_out/z_pw_bench.json 2> gpurun_out/z_pw_bench.err
timeout -k 10 200 python tools/kitti_eval.py none 3 > gpurun_out/z_kitti_eval.json 2> gpurun_out/z_kitti_eval.err
tail -2 gpurun_out/z_pytest.log; tail -2 gpurun_out/z_smoke.log
python - <<'PY'
import json
for f in ('z_bench_ref', 'z_bench_n1'):
    try:
        l = json.loads(open('gpurun_out/%s.json' % f).read().strip().splitlines()[-1])
        print(f, {k: l.get(k) for k in ('value', 'ms_per_step', 'parity_checked', 'gpu_launches')}, (l.get('roofline') or {}).get('frac'), l['e2e'], l.get('clocks'))
        for k, v in (l.get('extras') or {}).items():
            print(' ', k, json.dumps(v)[:260])
    except Exception as e:
        print(f, 'parse failed', e)
PY
cat gpurun_out/z_pw_bench.json
