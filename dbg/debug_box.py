import sys; sys.path.insert(0, '.')
import numpy as np
import vpp_b200 as vpp
from vpp_b200 import capi
from tests import oracle as orc
capi.check(capi.lib.vppb_init(0))
for shape in [(270,480),(64,341),(1080,1920)]:
    src = np.random.default_rng(1).integers(0,256,shape+(3,),dtype=np.uint8)
    S = vpp.Image2d.from_host(src, "vuchar3", border=2); vpp.fill_border_mirror(S)
    D = vpp.Image2d(*shape,"vuchar3"); vpp.fill(D, 0)
    vpp.box5x5(S, D)
    hs = orc.HostImage(*shape, "vuchar3", border=2, data=src, fill_border="mirror"); hd = orc.HostImage(*shape, "vuchar3")
    orc.load().vo_box5x5_u8(hs.ptr(), hd.ptr(), 3)
    got, exp = D.download().reshape(shape[0], -1), hd.get().reshape(shape[0], -1)
    bad = got != exp
    print(shape, "mismatch", bad.sum(), "of", bad.size)
    if bad.any():
        rows = np.where(bad.any(axis=1))[0]; cols = np.where(bad.any(axis=0))[0]
        print(" rows", rows[:10], "...", rows[-5:], " cols", cols[:12], "...", cols[-5:])
        r, c = np.argwhere(bad)[0]
        print(" first", r, c, "got", got[r, c:c+8], "exp", exp[r, c:c+8], "diff hist", np.unique((got.astype(int)-exp)[bad], return_counts=True))
