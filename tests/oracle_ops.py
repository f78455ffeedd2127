"""Composite reference operators assembled from the oracle's C functions (test infrastructure)."""
import ctypes as C

import numpy as np

from . import oracle as orc


def oracle_pyramid(img_u8_or_host, nlevels, pixel="u8", border=0, lib=None):
    """pyramid2d<V>(img, nlevels, 2, _border=) (pyramid.hh:146-198) -> list of HostImage."""
    o = lib or orc.load()
    if isinstance(img_u8_or_host, orc.HostImage):
        src = img_u8_or_host
        nr, nc = src.nrows, src.ncols
    else:
        a = np.asarray(img_u8_or_host)
        nr, nc = a.shape[:2]
        src = orc.HostImage(nr, nc, pixel, data=a)
    levels = []
    for _ in range(nlevels):
        levels.append(orc.HostImage(nr, nc, pixel, border=border))
        nr, nc = int(1 + nr / 2), int(1 + nc / 2)
    o.vo_copy(src.ptr(), levels[0].ptr(), 0)
    propagate(levels, pixel, o)
    return levels


def propagate(levels, pixel, o):
    kind = {"u8": 0, "vint2": 1, "vfloat2": 2}[pixel]
    o.vo_fill_border_mirror(levels[0].ptr())
    for i in range(1, len(levels)):
        o.vo_lowpass_sub2(levels[i - 1].ptr(), levels[i].ptr(), kind)
        o.vo_fill_border_mirror(levels[i].ptr())


def oracle_grad_pyramid(prev_levels, grad_pixel, border, o):
    g = [orc.HostImage(l.nrows, l.ncols, grad_pixel, border=border) for l in prev_levels]
    o.vo_scharr_u8(prev_levels[0].ptr(), g[0].ptr(), 1 if grad_pixel == "vfloat2" else 0)
    propagate(g, grad_pixel, o)
    return g


def oracle_lk(prev, nxt, grad, params, pts, prediction=None, lib=None):
    o = lib or orc.load()
    kp = np.ascontiguousarray(pts, dtype=np.float32).reshape(-1, 2)
    n = len(kp)
    flow = np.zeros((n, 2), dtype=np.float32)
    err = np.zeros(n, dtype=np.float32)
    pred = None
    if prediction is not None:
        pred = np.ascontiguousarray(prediction, dtype=np.float32).reshape(-1, 2)
    o.vo_lk_match_u8(orc.desc_array(prev), orc.desc_array(nxt), orc.desc_array(grad), C.byref(params), kp.ctypes.data,
                     pred.ctypes.data if pred is not None else None, n, flow.ctypes.data, err.ctypes.data)
    return flow, err


def oracle_lucas_kanade(i1, i2, pts, niterations=21, winsize=11, nscales=3, min_ev=0.0001, delta=0.1, prediction=None, lib=None):
    """lucas_kanade(i1, i2, ...) of lucas_kanade.hpp:135-184 on the oracle."""
    o = lib or orc.load()
    border = winsize // 2
    prev = oracle_pyramid(i1, nscales, "u8", border, o)
    nxt = oracle_pyramid(i2, nscales, "u8", border, o)
    grad = oracle_grad_pyramid(prev, "vint2", border, o)
    P = orc.VoLkParams(nlevels=nscales, min_scale=0, winsize=winsize, max_iter=niterations, grad_is_float=0, err_mode=0,
                       gate_on_max_err=0, min_ev=float(int(min_ev)), delta=float(int(delta)), max_err=0.0, factor=2.0,
                       pred_div=float(2 ** nscales))
    return oracle_lk(prev, nxt, grad, P, pts, prediction, o)
