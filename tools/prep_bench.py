import sys, ctypes as C, numpy as np, time
sys.path.insert(0, '.')
import vpp_b200 as vpp
from vpp_b200 import capi
from tests import scenes
capi.check(capi.lib.vppb_init(0))
f1, f2, pts = scenes.lk_pair(1080, 1920, 100, seed=5)
I1, I2 = vpp.Image2d.from_host(f1, "u8"), vpp.Image2d.from_host(f2, "u8")
prev, nxt = vpp.Pyramid2d((1080, 1920), 3, 2, pixel="u8", border=4), vpp.Pyramid2d((1080, 1920), 3, 2, pixel="u8", border=4)
grad = vpp.Pyramid2d((1080, 1920), 3, 2, pixel="vfloat2", border=4)
for _ in range(5):
    vpp.pyrlk_prepare(I1, I2, prev, nxt, grad)
capi.check(capi.lib.vppb_sync(None))
t0 = time.perf_counter()
for _ in range(200):
    vpp.pyrlk_prepare(I1, I2, prev, nxt, grad)
capi.check(capi.lib.vppb_sync(None))
print("ms per prepare (wall, 200 calls):", (time.perf_counter() - t0) / 200 * 1e3)
