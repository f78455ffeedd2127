import sys, ctypes as C; sys.path.insert(0, '.')
import numpy as np, torch
import vpp_b200 as vpp
from vpp_b200 import capi
from vpp_b200.ops import _DeviceBuffer
from tests import scenes
capi.check(capi.lib.vppb_init(0))
H, W = 2160, 3840
g = scenes.rectangles_scene(H, W, seed=42)
U = vpp.Image2d.from_host(g, "u8", border=3); vpp.fill_border_mirror(U)
ws = _DeviceBuffer(capi.lib.vppb_fast9_workspace_bytes(H, W, 10)); cap = H * W // 8
kp, sc, cnt = _DeviceBuffer(cap * 8), _DeviceBuffer(cap * 4), C.c_int32()
for mode in (0, 1, 2):
    for _ in range(3):
        capi.check(capi.lib.vppb_fast9_u8(U.ptr(), 20, None, mode, 10, 0, ws.ptr, ws.nbytes, kp.ptr, sc.ptr, cap, C.byref(cnt), None))
    print("mode", mode, "kps", cnt.value)
