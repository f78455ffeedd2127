#!/bin/bash
# round 2, GPU call D: parity, C++ lambda lowering bench, FAST9 (warp-autonomous band kernel) timing + ncu
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests -x -q -m gpu > gpurun_out/d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/d_pytest.log
timeout 120 tests/cpp/_build/pw_bench > gpurun_out/d_pw_bench.json 2> gpurun_out/d_pw_bench.err
timeout -k 10 600 python tools/fast_bench.py > gpurun_out/d_fast_bench.json 2> gpurun_out/d_fast_bench.err
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
timeout -k 10 300 ncu --set full --clock-control none --import-source on -k regex:k_fast9_band -s 1 -c 1 -f -o gpurun_out/d_prof_fast4k python /tmp/fast4k.py > gpurun_out/d_ncu_fast.log 2>&1
timeout -k 10 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/d_fast_launches.csv python /tmp/fast4k.py > /dev/null 2>&1
tail -3 gpurun_out/d_pytest.log
cat gpurun_out/d_pw_bench.json
tail -4 gpurun_out/d_fast_bench.err
