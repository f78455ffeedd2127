#!/bin/bash
# round 2, GPU call M: launch list of one semi-dense flow call at 1080p (where do the 8 ms go?)
mkdir -p gpurun_out
cat > /tmp/sdof1.py <<'PY'
import sys, ctypes as C, numpy as np
sys.path.insert(0, '.')
import vpp_b200 as vpp
from vpp_b200 import capi
from vpp_b200.ops import _DeviceBuffer
from tests import scenes
capi.check(capi.lib.vppb_init(0))
g1, g2, _ = scenes.lk_pair(1080, 1920, 4, seed=55, shift=(3.0, -2.0), margin=10)
G = vpp.Image2d.from_host(g1, "u8", border=3); vpp.fill_border_mirror(G)
kps = vpp.fast9(G, 10, blockwise=True, block_size=10)
n = len(kps)
P = capi.VppbSdofParams(9, 3, 0, 2, 5)
I1, I2 = vpp.Image2d.from_host(g1, "u8"), vpp.Image2d.from_host(g2, "u8")
p1, p2 = vpp.Pyramid2d(I1, 3, 2, border=18), vpp.Pyramid2d(I2, 3, 2, border=18)
wsb = _DeviceBuffer(capi.lib.vppb_sdof_workspace_bytes(1080, 1920, C.byref(P)))
d_kp = _DeviceBuffer(kps.nbytes).from_host(kps)
d_pos, d_dist, d_valid = _DeviceBuffer(n * 8), _DeviceBuffer(n * 4), _DeviceBuffer(n)
for _ in range(3):
    capi.check(capi.lib.vppb_sdof_u8(p1.desc_array(), p2.desc_array(), C.byref(P), d_kp.ptr, n, wsb.ptr, wsb.nbytes, d_pos.ptr, d_dist.ptr, d_valid.ptr, None))
capi.check(capi.lib.vppb_sync(None))
print(n)
PY
timeout -k 10 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -c 200 --csv --log-file gpurun_out/m_sdof_launches.csv python /tmp/sdof1.py > /dev/null 2>&1
grep k_sdof gpurun_out/m_sdof_launches.csv | awk -F'","' '{print $5, $(NF-2), $NF}' | tail -44
