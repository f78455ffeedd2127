#!/bin/bash
# round 2, GPU call A: parity (all -m gpu tests incl. BASELINE sizes), box A/B timing, ncu of the streaming box kernel
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/a_smi.txt 2>&1
timeout -k 10 600 python -m pytest tests -x -q -m gpu > gpurun_out/a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/a_pytest.log
timeout -k 10 900 python tools/box_bench.py > gpurun_out/a_box_bench.json 2> gpurun_out/a_box_bench.err
cat > /tmp/box4k.py <<'PY'
import sys, numpy as np
sys.path.insert(0, '.')
import vpp_b200 as vpp
from vpp_b200 import capi
capi.check(capi.lib.vppb_init(0))
r = np.random.default_rng(1)
H, W = 2160, 3840
S = vpp.Image2d.from_host(r.integers(0, 256, (H, W, 3), dtype=np.uint8), "vuchar3", border=2); vpp.fill_border_mirror(S)
D = vpp.Image2d(H, W, "vuchar3")
for _ in range(4): vpp.box5x5(S, D)
capi.check(capi.lib.vppb_sync(None))
PY
for lw in 4 8; do
VPPB_BOX_LW=$lw timeout -k 10 300 ncu --set full --clock-control none --import-source on -k regex:k_box5_stream -s 2 -c 1 -f -o gpurun_out/a_prof_box4k_lw$lw python /tmp/box4k.py > gpurun_out/a_ncu_lw$lw.log 2>&1
done
tail -3 gpurun_out/a_pytest.log
cat gpurun_out/a_box_bench.err | tail -12
