import sys, ctypes as C; sys.path.insert(0, '.')
import numpy as np, torch
import vpp_b200 as vpp
from vpp_b200 import capi
from vpp_b200.ops import _DeviceBuffer
from tests import scenes
capi.check(capi.lib.vppb_init(0))
f1, f2, pts = scenes.lk_pair(1080, 1920, 10000, seed=5)
I1, I2 = vpp.Image2d.from_host(f1, "u8"), vpp.Image2d.from_host(f2, "u8")
prev, nxt = vpp.Pyramid2d(I1, 3, 2, border=3), vpp.Pyramid2d(I2, 3, 2, border=3)
grad = vpp.Pyramid2d((1080, 1920), 3, 2, pixel="vint2", border=3)
vpp.scharr(prev[0], grad[0]); grad.propagate_level0()
d_kp = _DeviceBuffer(pts.nbytes).from_host(pts)
d_flow, d_err = _DeviceBuffer(len(pts) * 8), _DeviceBuffer(len(pts) * 4)
P = capi.VppbLkParams(nlevels=3, min_scale=0, winsize=7, max_iter=21, grad_is_float=0, err_mode=0, gate_on_max_err=0, min_ev=0.0, delta=0.0, max_err=0.0, factor=2.0, pred_div=8.0)
pa, na, ga = prev.desc_array(), nxt.desc_array(), grad.desc_array()
for _ in range(3):
    capi.check(capi.lib.vppb_lk_match_u8(pa, na, ga, C.byref(P), d_kp.ptr, None, len(pts), d_flow.ptr, d_err.ptr, None))
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10):
    capi.check(capi.lib.vppb_lk_match_u8(pa, na, ga, C.byref(P), d_kp.ptr, None, len(pts), d_flow.ptr, d_err.ptr, None))
b.record(); torch.cuda.synchronize()
print("lk ms", a.elapsed_time(b) / 10)
