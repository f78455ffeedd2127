#!/usr/bin/env python
"""FAST9 timing at 4K (GPU box only): device time of the queued work (CUDA events around vppb_fast9_u8_async) and the
wall time of the Python fast9() call, band kernel vs the round-1 warp-per-32-pixels kernel (VPPB_FAST_IMPL=warp)."""
import ctypes as C
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child():
    import numpy as np
    import torch

    import vpp_b200 as vpp
    from vpp_b200 import capi, ops
    from tests import scenes

    capi.check(capi.lib.vppb_init(0))
    stream = torch.cuda.current_stream()
    sp = C.c_void_p(stream.cuda_stream)
    out = {}
    for (H, W, tag) in [(2160, 3840, "4k"), (1080, 1920, "1080p")]:
        img = scenes.rectangles_scene(H, W, seed=42)
        G = vpp.Image2d.from_host(img, "u8", border=3)
        vpp.fill_border_mirror(G)
        mask = np.full(img.shape, 1, dtype=np.uint8)
        M = vpp.Image2d.from_host(mask, "u8")
        for mode, name, m in ((capi.FAST_ALL, "all", None), (capi.FAST_BLOCKWISE, "blockwise_mask1", M), (capi.FAST_LOCAL_MAXIMA, "local_maxima", None)):
            cap = H * W // 8
            ent = ops._fast_buffers(G, 10, cap, True)

            def run():
                capi.check(capi.lib.vppb_fast9_u8_async(G.ptr(), 20, m.ptr() if m is not None else None, mode, 10, 0, ent["ws"].ptr, ent["ws"].nbytes,
                                                         ent["kps"].ptr, None, cap, ent["count"].ptr, sp))
            for _ in range(5):
                run()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            for _ in range(50):
                run()
            b.record(stream)
            torch.cuda.synchronize()
            n = int(ent["count"].to_host(np.int32, 1, sp)[0])
            us = a.elapsed_time(b) * 1e3 / 50
            out["%s_%s" % (tag, name)] = {"device_us": round(us, 2), "kps": n, "GBps": round((H * W + 8 * n) / us / 1e3, 1)}
        t0 = time.perf_counter()
        for _ in range(20):
            k = vpp.fast9(G, 20, stream=sp)
        out[tag + "_python_fast9_wall_us"] = round((time.perf_counter() - t0) / 20 * 1e6, 1)
        out[tag + "_kps"] = len(k)
    print(json.dumps(out))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child()
        sys.exit(0)
    res = {}
    for name, env in (("band", {}), ("warp", {"VPPB_FAST_IMPL": "warp"})):
        e = dict(os.environ)
        e.update(env)
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=e, capture_output=True, text=True, timeout=300)
        res[name] = json.loads(r.stdout.strip().splitlines()[-1]) if r.returncode == 0 else {"error": (r.stderr or r.stdout)[-600:]}
        print(name, json.dumps(res[name]), file=sys.stderr)
    print(json.dumps(res, indent=1))
