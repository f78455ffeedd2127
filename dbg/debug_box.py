import sys; sys.path.insert(0, '.')
import numpy as np
import vpp_b200 as vpp
from vpp_b200 import capi
capi.check(capi.lib.vppb_init(0))
src = np.random.default_rng(1).integers(0,256,(270,480,3),dtype=np.uint8)
S = vpp.Image2d.from_host(src, "vuchar3", border=2); vpp.fill_border_mirror(S)
D = vpp.Image2d(270,480,"vuchar3")
vpp.box5x5(S, D)
print(capi.lib.vppb_sync(None), capi.lib.vppb_last_error())
print(D.download()[:2,:4])
