#!/usr/bin/env python
"""pyrLK at BASELINE config 4 (1080p, 3 levels, 10 000 keypoints, 7x7, vfloat2 gradient): CUDA-event time of vppb_lk_match_u8 and of
vppb_pyrlk_prepare, parity against the oracle.  Usage (GPU box): python tools/lk_bench.py"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vpp_b200 as vpp  # noqa: E402
from vpp_b200 import capi  # noqa: E402
from vpp_b200.ops import _DeviceBuffer  # noqa: E402
from tests import oracle as orc, scenes  # noqa: E402
from tests.oracle_ops import oracle_grad_pyramid, oracle_lk, oracle_pyramid  # noqa: E402

capi.check(capi.lib.vppb_init(0))
torch.cuda.set_device(0)
stream = torch.cuda.current_stream()
sp = C.c_void_p(stream.cuda_stream)
f1, f2, pts = scenes.lk_pair(1080, 1920, 10000, seed=5)
I1, I2 = vpp.Image2d.from_host(f1, "u8"), vpp.Image2d.from_host(f2, "u8")
prev, nxt = vpp.Pyramid2d((1080, 1920), 3, 2, pixel="u8", border=4), vpp.Pyramid2d((1080, 1920), 3, 2, pixel="u8", border=4)
grad = vpp.Pyramid2d((1080, 1920), 3, 2, pixel="vfloat2", border=4)
d_kp = _DeviceBuffer(pts.nbytes).from_host(pts)
d_flow, d_err = _DeviceBuffer(len(pts) * 8), _DeviceBuffer(len(pts) * 4)
P = capi.VppbLkParams(nlevels=3, min_scale=0, winsize=7, max_iter=21, grad_is_float=1, err_mode=capi.LK_ERR_SAD_OVER_MAD, gate_on_max_err=1, min_ev=0.01,
                      delta=0.01, max_err=0.6, factor=2.0, pred_div=1.0)
pa, na, ga = prev.desc_array(), nxt.desc_array(), grad.desc_array()


def build():
    vpp.pyrlk_prepare(I1, I2, prev, nxt, grad, sp)


def lk():
    capi.check(capi.lib.vppb_lk_match_u8(pa, na, ga, C.byref(P), d_kp.ptr, None, len(pts), d_flow.ptr, d_err.ptr, sp))


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(reps):
        fn()
    b.record(stream)
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


build(); lk()
flow, err = d_flow.to_host(np.float32, len(pts) * 2, sp).reshape(-1, 2), d_err.to_host(np.float32, len(pts), sp)
omp = orc.load(omp=True)
rprev, rnxt = oracle_pyramid(f1, 3, "u8", 4, omp), oracle_pyramid(f2, 3, "u8", 4, omp)
rgrad = oracle_grad_pyramid(rprev, "vfloat2", 4, omp)
RP = orc.VoLkParams(nlevels=3, min_scale=0, winsize=7, max_iter=21, grad_is_float=1, err_mode=1, gate_on_max_err=1, min_ev=0.01, delta=0.01, max_err=0.6, factor=2.0, pred_div=1.0)
rflow, rerr = oracle_lk(rprev, rnxt, rgrad, RP, pts, lib=omp)
ok = bool(np.array_equal(err >= 3e38, rerr >= 3e38) and np.array_equal(flow.view(np.int32), rflow.view(np.int32)))
ms_build = min(timed(build, 20) for _ in range(3))
for w_ in ("4", "2", "1", "4", "1"):  # warps per CTA of the matching kernel (read per call), A/B in one process
    os.environ["VPPB_LK_WARPS"] = w_
    lk()
    f_ = d_flow.to_host(np.float32, len(pts) * 2, sp).reshape(-1, 2)
    print("VPPB_LK_WARPS=%s ms_match %.4f bit_exact %s" % (w_, min(timed(lk, 20) for _ in range(3)), bool(np.array_equal(f_.view(np.int32), rflow.view(np.int32)))))
os.environ.pop("VPPB_LK_WARPS", None)
ms_lk = min(timed(lk, 20) for _ in range(3))
print({"ms_match": ms_lk, "ms_prepare": ms_build, "kpts_per_s": len(pts) / ((ms_lk + ms_build) / 1e3), "kpts_per_s_match_only": len(pts) / (ms_lk / 1e3), "bit_exact": ok,
       "env": {k: v for k, v in os.environ.items() if k.startswith("VPPB_")}})
