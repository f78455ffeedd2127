#!/usr/bin/env python
"""Time the semi-dense optical flow (vppb_sdof_u8) at the sizes of BASELINE config 5 (8K) and 1080p.
Usage (GPU box): python tools/sdof_bench.py"""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vpp_b200 as vpp  # noqa: E402
from vpp_b200 import capi  # noqa: E402
from vpp_b200.ops import _DeviceBuffer  # noqa: E402
from tests import scenes  # noqa: E402

capi.check(capi.lib.vppb_init(0))
for (H, W) in [(1080, 1920), (4320, 7680)]:
    base = scenes.rectangles_scene(H + 16, W + 16, seed=5, noise=3)
    f1, f2 = base[8:8 + H, 8:8 + W].copy(), base[5:5 + H, 10:10 + W].copy()  # motion (3, -2)
    G = vpp.Image2d.from_host(f1, "u8", border=3)
    vpp.fill_border_mirror(G)
    kps = vpp.fast9(G, 10, blockwise=True, block_size=10)
    n = len(kps)
    P = capi.VppbSdofParams(9, 3, 0, 2, 5)  # video_extruder's settings
    p1 = vpp.Pyramid2d(vpp.Image2d.from_host(f1, "u8"), 3, 2, border=18)
    p2 = vpp.Pyramid2d(vpp.Image2d.from_host(f2, "u8"), 3, 2, border=18)
    ws = _DeviceBuffer(capi.lib.vppb_sdof_workspace_bytes(H, W, C.byref(P)))
    d_kp = _DeviceBuffer(kps.nbytes).from_host(kps)
    d_pos, d_dist, d_valid = _DeviceBuffer(n * 8), _DeviceBuffer(n * 4), _DeviceBuffer(n)

    def run():
        capi.check(capi.lib.vppb_sdof_u8(p1.desc_array(), p2.desc_array(), C.byref(P), d_kp.ptr, n, ws.ptr, ws.nbytes, d_pos.ptr, d_dist.ptr, d_valid.ptr, None))
        capi.check(capi.lib.vppb_sync(None))

    for sched in ("levels", "dataflow", None):  # the opt-in schedules first, the default (one cooperative launch, relaxation) last: its numbers print below
        if sched:
            os.environ["VPPB_SDOF_SCHEDULE"] = sched
        else:
            os.environ.pop("VPPB_SDOF_SCHEDULE", None)
        run()
        t0 = time.perf_counter()
        reps = 3
        for _ in range(reps):
            run()
        dt = (time.perf_counter() - t0) / reps
        if sched:
            print("%dx%d: schedule=%s %.2f ms per frame pair" % (W, H, sched, dt * 1e3))
    pos = d_pos.to_host(np.int32, n * 2).reshape(-1, 2)
    valid = d_valid.to_host(np.uint8, n).astype(bool)
    flow = (pos - kps)[valid]
    ok = (np.abs(flow - np.array([3, -2])).max(axis=1) <= 1).mean() if valid.any() else 0.0
    print("%dx%d: %d keypoints, %d reported, %.1f ms per frame pair (flow only, pyramids prebuilt) -> %.2f M kps/s; %.0f %% within 1 px of the true motion"
          % (W, H, n, valid.sum(), dt * 1e3, n / dt / 1e6, 100 * ok))
