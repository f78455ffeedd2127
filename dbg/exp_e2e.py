"""PCIe legs of the e2e path: linear vs 2D copies, duplex overlap, pipeline depth."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import vpp_b200 as vpp
from vpp_b200 import capi

dev = torch.device("cuda:0"); capi.check(capi.lib.vppb_init(0))
H, W = 1080, 1920; rowb = W * 3; NF = 64
hin = [torch.randint(0, 255, (H, W, 3), dtype=torch.uint8).pin_memory() for _ in range(8)]
hout = [torch.empty((H, W, 3), dtype=torch.uint8).pin_memory() for _ in range(8)]
lin = [torch.empty(H * rowb, dtype=torch.uint8, device=dev) for _ in range(8)]

def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    return best

def report(name, t, nbytes):
    print("%-38s %8.1f us/frame %7.1f GB/s" % (name, t / NF * 1e6, nbytes / t / 1e9), flush=True)

s = [torch.cuda.Stream() for _ in range(8)]
def h2d_lin():
    with torch.cuda.stream(s[0]):
        for i in range(NF): lin[i % 8].copy_(hin[i % 8].view(-1), non_blocking=True)
report("H2D linear 1 stream", timeit(h2d_lin), NF * H * rowb)
def d2h_lin():
    with torch.cuda.stream(s[1]):
        for i in range(NF): hout[i % 8].view(-1).copy_(lin[i % 8], non_blocking=True)
report("D2H linear 1 stream", timeit(d2h_lin), NF * H * rowb)
def duplex():
    h2d_lin(); d2h_lin()
report("H2D+D2H linear duplex (bytes each way)", timeit(duplex), NF * H * rowb)
src = [vpp.Image2d(H, W, "vuchar3", border=2) for _ in range(8)]
dst = [vpp.Image2d(H, W, "vuchar3") for _ in range(8)]
print("src pitch", src[0].pitch, "dst pitch", dst[0].pitch)
def h2d_2d():
    st = C.c_void_p(s[0].cuda_stream)
    for i in range(NF): capi.check(capi.lib.vppb_upload(src[i % 8].ptr(), C.c_void_p(hin[i % 8].data_ptr()), rowb, 0, st))
report("H2D 2D (vppb_upload) 1 stream", timeit(h2d_2d), NF * H * rowb)
def d2h_2d():
    st = C.c_void_p(s[1].cuda_stream)
    for i in range(NF): capi.check(capi.lib.vppb_download(dst[i % 8].ptr(), C.c_void_p(hout[i % 8].data_ptr()), rowb, 0, st))
report("D2H (vppb_download) 1 stream", timeit(d2h_2d), NF * H * rowb)
def h2d_2d_2s():
    for i in range(NF):
        st = C.c_void_p(s[i & 1].cuda_stream)
        capi.check(capi.lib.vppb_upload(src[i % 8].ptr(), C.c_void_p(hin[i % 8].data_ptr()), rowb, 0, st))
report("H2D 2D 2 streams", timeit(h2d_2d_2s), NF * H * rowb)

for ns in (1, 2, 3, 4, 6, 8):
    def e2e():
        for i in range(NF):
            k = i % ns
            st = C.c_void_p(s[k].cuda_stream)
            capi.check(capi.lib.vppb_upload(src[k].ptr(), C.c_void_p(hin[i % 8].data_ptr()), rowb, 0, st))
            capi.check(capi.lib.vppb_fill_border_mirror(src[k].ptr(), st))
            capi.check(capi.lib.vppb_box5x5_u8c3(src[k].ptr(), dst[k].ptr(), st))
            capi.check(capi.lib.vppb_download(dst[k].ptr(), C.c_void_p(hout[i % 8].data_ptr()), rowb, 0, st))
    t = timeit(e2e)
    print("e2e %d streams: %7.1f us/frame  %6.2f Gpix/s" % (ns, t / NF * 1e6, NF * H * W / t / 1e9), flush=True)

# three-stage pipeline: dedicated upload / compute / download streams linked by events
up, cp, dn = s[0], s[1], s[2]
for depth in (2, 3, 4, 8):
    ev_up = [torch.cuda.Event() for _ in range(depth)]; ev_cp = [torch.cuda.Event() for _ in range(depth)]; ev_dn = [torch.cuda.Event() for _ in range(depth)]
    def pipe():
        for i in range(NF):
            k = i % depth
            if i >= depth: up.wait_event(ev_cp[k])          # src[k] consumed by the previous box on it
            capi.check(capi.lib.vppb_upload(src[k].ptr(), C.c_void_p(hin[i % 8].data_ptr()), rowb, 0, C.c_void_p(up.cuda_stream)))
            ev_up[k].record(up)
            cp.wait_event(ev_up[k])
            if i >= depth: cp.wait_event(ev_dn[k])          # dst[k] drained
            capi.check(capi.lib.vppb_fill_border_mirror(src[k].ptr(), C.c_void_p(cp.cuda_stream)))
            capi.check(capi.lib.vppb_box5x5_u8c3(src[k].ptr(), dst[k].ptr(), C.c_void_p(cp.cuda_stream)))
            ev_cp[k].record(cp)
            dn.wait_event(ev_cp[k])
            capi.check(capi.lib.vppb_download(dst[k].ptr(), C.c_void_p(hout[i % 8].data_ptr()), rowb, 0, C.c_void_p(dn.cuda_stream)))
            ev_dn[k].record(dn)
    t = timeit(pipe)
    print("3-stage pipeline depth %d: %7.1f us/frame  %6.2f Gpix/s" % (depth, t / NF * 1e6, NF * H * W / t / 1e9), flush=True)

# staged upload: linear H2D into a tight device buffer, then copy+mirror into the bordered image in one launch
stage = [torch.empty(H * rowb, dtype=torch.uint8, device=dev) for _ in range(8)]
sdesc = [capi.VppbImg(base=t.data_ptr(), alloc=None, nrows=H, ncols=W, pitch=rowb, border=0, elem_bytes=3, align=16) for t in stage]
for ns in (2, 3, 4):
    def e2e_staged():
        for i in range(NF):
            k = i % ns
            st = C.c_void_p(s[k].cuda_stream)
            with torch.cuda.stream(s[k]):
                stage[k].copy_(hin[i % 8].view(-1), non_blocking=True)
            capi.check(capi.lib.vppb_copy2d_mirror(C.byref(sdesc[k]), src[k].ptr(), st))
            capi.check(capi.lib.vppb_box5x5_u8c3(src[k].ptr(), dst[k].ptr(), st))
            capi.check(capi.lib.vppb_download(dst[k].ptr(), C.c_void_p(hout[i % 8].data_ptr()), rowb, 0, st))
    t = timeit(e2e_staged)
    print("e2e staged-linear %d streams: %7.1f us/frame  %6.2f Gpix/s" % (ns, t / NF * 1e6, NF * H * W / t / 1e9), flush=True)
# the staged result must equal the direct path's
capi.check(capi.lib.vppb_upload(src[0].ptr(), C.c_void_p(hin[0].data_ptr()), rowb, 0, None)); vpp.fill_border_mirror(src[0]); vpp.box5x5(src[0], dst[0])
a = dst[0].download()
capi.check(capi.lib.vppb_sync(None))
stage[1].copy_(hin[0].view(-1)); capi.check(capi.lib.vppb_copy2d_mirror(C.byref(sdesc[1]), src[1].ptr(), None)); vpp.box5x5(src[1], dst[1])
print("staged == direct:", np.array_equal(a, dst[1].download()))
