#!/usr/bin/env python
"""Per-kernel timing of every C-ABI entry point at its BASELINE size: CUDA events on the launch stream,
a rotating set of images larger than L2, achieved algorithmic GB/s against MEASURED_PEAKS.json.
Usage (GPU box): python tools/kernel_bench.py > gpurun_out/kernel_bench.json"""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import vpp_b200 as vpp  # noqa: E402
from vpp_b200 import capi  # noqa: E402
from tests import scenes  # noqa: E402

capi.check(capi.lib.vppb_init(0))
PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
L2 = 140e6
stream = torch.cuda.current_stream()
sp = C.c_void_p(stream.cuda_stream)
out = {}


def timed(fn, nset, reps=20):
    for i in range(nset):
        fn(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for r in range(reps):
        for i in range(nset):
            fn(i)
    b.record(stream)
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / (reps * nset)  # us per call


def report(name, us, alg_bytes, note=""):
    gbs = alg_bytes / (us * 1e-6) / 1e9
    out[name] = {"us": round(us, 2), "alg_MB": round(alg_bytes / 1e6, 2), "GBps": round(gbs, 1), "frac_of_peak": round(gbs / PEAK, 3), "note": note}
    print("%-34s %9.2f us  %8.2f MB  %8.1f GB/s  %5.1f %%  %s" % (name, us, alg_bytes / 1e6, gbs, 100 * gbs / PEAK, note), file=sys.stderr)


def nsets(bytes_per_set):
    return max(2, int(np.ceil(L2 / bytes_per_set)))


rng = np.random.default_rng(0)
for (H, W, tag) in [(1080, 1920, "1080p"), (2160, 3840, "4k")]:
    # --- int32 add / fill / copy
    n = nsets(12 * H * W)
    A = [vpp.Image2d(H, W, "i32") for _ in range(n)]
    B = [vpp.Image2d(H, W, "i32") for _ in range(n)]
    Cc = [vpp.Image2d(H, W, "i32") for _ in range(n)]
    v = np.array([7], np.int32)
    for x in B + Cc:
        capi.check(capi.lib.vppb_fill(x.ptr(), v.ctypes.data, 0, sp))
    report("add_i32_" + tag, timed(lambda i: capi.lib.vppb_pw_add_i32(A[i].ptr(), B[i].ptr(), Cc[i].ptr(), sp), n), 12.0 * H * W)
    report("fill_i32_" + tag, timed(lambda i: capi.lib.vppb_fill(A[i].ptr(), v.ctypes.data, 0, sp), n), 4.0 * H * W)
    report("copy_i32_" + tag, timed(lambda i: capi.lib.vppb_copy2d(B[i].ptr(), A[i].ptr(), 0, sp), n), 8.0 * H * W)
    del A, B, Cc
    # --- vuchar3 box + mirror border
    n = nsets(6 * H * W)
    f = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    S = [vpp.Image2d.from_host(f, "vuchar3", border=2) for _ in range(n)]
    D = [vpp.Image2d(H, W, "vuchar3") for _ in range(n)]
    report("fill_border_mirror_u8c3_b2_" + tag, timed(lambda i: capi.lib.vppb_fill_border_mirror(S[i].ptr(), sp), n), 2.0 * 3 * 2 * (2 * (W + 4) + 2 * H), "border px only")
    report("box5x5_u8c3_" + tag, timed(lambda i: capi.lib.vppb_box5x5_u8c3(S[i].ptr(), D[i].ptr(), sp), n), 6.0 * H * W)
    nb = min(n, 32)
    bi, bo = (capi.VppbImg * nb)(*[x.desc for x in S[:nb]]), (capi.VppbImg * nb)(*[x.desc for x in D[:nb]])
    us_b = timed(lambda i: capi.lib.vppb_box5x5_u8c3_batch(bi, bo, nb, sp) if i == 0 else 0, nb) * 1.0  # one call per sweep of the set
    report("box5x5_u8c3_batch%d_per_frame_" % nb + tag, us_b, 6.0 * H * W, "one persistent launch over %d frames; time per frame" % nb)
    del S, D
    # --- u8: box, scharr, pyramid step, fast9 pieces
    n = nsets(9 * H * W)
    g = scenes.rectangles_scene(H, W, seed=42)
    U = [vpp.Image2d.from_host(g, "u8", border=3) for _ in range(n)]
    for x in U:
        vpp.fill_border_mirror(x)
    U2 = [vpp.Image2d(H, W, "u8") for _ in range(n)]
    report("box5x5_u8_" + tag, timed(lambda i: capi.lib.vppb_box5x5_u8(U[i].ptr(), U2[i].ptr(), sp), n), 2.0 * H * W)
    Gi = [vpp.Image2d(H, W, "vint2", border=2) for _ in range(n)]
    report("scharr_u8_vint2_" + tag, timed(lambda i: capi.lib.vppb_scharr_u8(U[i].ptr(), Gi[i].ptr(), 0, sp), n), 9.0 * H * W)
    h2, w2 = int(1 + H / 2), int(1 + W / 2)
    L1 = [vpp.Image2d(h2, w2, "u8", border=3) for _ in range(n)]
    report("lowpass_sub2_u8_" + tag, timed(lambda i: capi.lib.vppb_lowpass_sub2(U[i].ptr(), L1[i].ptr(), 0, sp), n), 1.0 * H * W + h2 * w2)
    for x in Gi:
        vpp.fill_border_mirror(x)
    G1 = [vpp.Image2d(h2, w2, "vint2", border=3) for _ in range(n)]
    report("lowpass_sub2_vint2_" + tag, timed(lambda i: capi.lib.vppb_lowpass_sub2(Gi[i].ptr(), G1[i].ptr(), 1, sp), n), 8.0 * H * W + 8.0 * h2 * w2)
    report("lowpass_sub2_mirror_u8_" + tag, timed(lambda i: capi.lib.vppb_lowpass_sub2_mirror(U[i].ptr(), L1[i].ptr(), 0, sp), n), 1.0 * H * W + h2 * w2)
    report("lowpass_sub2_mirror_vint2_" + tag, timed(lambda i: capi.lib.vppb_lowpass_sub2_mirror(Gi[i].ptr(), G1[i].ptr(), 1, sp), n), 8.0 * H * W + 8.0 * h2 * w2)
    report("scharr_u8_vint2_mirror_" + tag, timed(lambda i: capi.lib.vppb_scharr_u8_mirror(U[i].ptr(), Gi[i].ptr(), 0, sp), n), 9.0 * H * W)
    P3 = [vpp.Pyramid2d((H, W), 3, 2, pixel="u8", border=3) for _ in range(min(n, 4))]
    report("pyramid3_u8_update_" + tag, timed(lambda i: P3[i % len(P3)].update(U[i], sp), n), 2.0 * H * W + 1.25 * (H * W + h2 * w2), "copy+mirror, 2 x level+mirror: 3 launches")
    del Gi, G1, L1, U2, P3
    # fast9: whole call (detect + scan + emit + count read-back) through the C-ABI with preallocated buffers
    from vpp_b200.ops import _DeviceBuffer

    ws = _DeviceBuffer(capi.lib.vppb_fast9_workspace_bytes(H, W, 10))
    cap = H * W // 8
    kp, cnt = _DeviceBuffer(cap * 8), C.c_int32()

    def fast(i):
        capi.check(capi.lib.vppb_fast9_u8(U[i].ptr(), 20, None, 0, 10, 0, ws.ptr, ws.nbytes, kp.ptr, None, cap, C.byref(cnt), sp))

    us = timed(fast, n, reps=5)
    report("fast9_u8_" + tag, us, 1.0 * H * W + 8.0 * cnt.value, "%d kps, includes the host count read-back (sync)" % cnt.value)
    del U
print(json.dumps(out))
