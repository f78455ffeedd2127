"""Property-based pinning (hypothesis) of the oracle against the reference's own headers (oracle/_ref) over
random geometries: ragged sizes, borders, alignments, thresholds, masks.  CPU only."""
import ctypes as C
import os

import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from tests import oracle as orc
from tests.test_oracle_vs_ref import REF, _load

pytestmark = pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref not built (needs /root/reference)")
SET = dict(max_examples=1500, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow])


@pytest.fixture(scope="module")
def libs(built):
    return _load(REF), orc.load()


def _img(seed, nr, nc, pix, border, aligned, lo=0, hi=256):
    dt, ch = orc.PIXEL_TYPES[pix]
    d = np.random.default_rng(seed).integers(lo, hi, (nr, nc) + ((ch,) if ch > 1 else ())).astype(dt)
    return orc.HostImage(nr, nc, pix, border=border, aligned=aligned, data=d), d


@settings(**SET)
@given(nr=st.integers(1, 40), nc=st.integers(1, 60), border=st.integers(0, 6), aligned=st.sampled_from([1, 4, 16, 32, 128]),
       pix=st.sampled_from(["u8", "vuchar3", "i32", "vint2"]), seed=st.integers(0, 2 ** 16))
def test_border_fills_any_geometry(libs, nr, nc, border, aligned, pix, seed):
    ref, o = libs
    border = min(border, nr, nc)
    h1, d = _img(seed, nr, nc, pix, border, aligned)
    h2 = orc.HostImage(nr, nc, pix, border=border, aligned=aligned, data=d)
    ref.vppref_fill_border_mirror(h1.ptr()); o.vo_fill_border_mirror(h2.ptr())
    assert np.array_equal(h1.get(True), h2.get(True))
    if pix != "vint2":
        ref.vppref_fill_border_closest(h1.ptr()); o.vo_fill_border_closest(h2.ptr())
        assert np.array_equal(h1.get(True), h2.get(True))


@settings(**SET)
@given(nr=st.integers(1, 40), nc=st.integers(1, 50), aligned=st.sampled_from([1, 16, 32, 128]), seed=st.integers(0, 2 ** 16), kind=st.sampled_from(["i32", "vuchar3"]))
def test_box_any_geometry(libs, nr, nc, aligned, seed, kind):
    ref, o = libs
    b = min(2 + seed % 3, max(nr, 2), max(nc, 2))
    if nr < 2 or nc < 2:
        b = 2  # mirror needs border <= size; tiny images are filled by value instead
    hs, d = _img(seed, nr, nc, kind, b, aligned, -500 if kind == "i32" else 0, 1000 if kind == "i32" else 256)
    if b <= nr and b <= nc:
        o.vo_fill_border_mirror(hs.ptr())
    else:
        v = np.zeros(4, np.int32)
        o.vo_fill_border_value(hs.ptr(), v.ctypes.data)
    d1, d2 = orc.HostImage(nr, nc, kind, aligned=aligned), orc.HostImage(nr, nc, kind, aligned=aligned)
    if kind == "i32":
        ref.vppref_box5x5_i32(hs.ptr(), d1.ptr()); o.vo_box5x5_i32(hs.ptr(), d2.ptr())
    else:
        ref.vppref_box5x5_u8c3(hs.ptr(), d1.ptr()); o.vo_box5x5_u8(hs.ptr(), d2.ptr(), 3)
    assert np.array_equal(d1.get(), d2.get())


@settings(**SET)
@given(nr=st.integers(3, 45), nc=st.integers(3, 60), seed=st.integers(0, 2 ** 16), as_float=st.booleans())
def test_scharr_and_lowpass_any_geometry(libs, nr, nc, seed, as_float):
    ref, o = libs
    h, _ = _img(seed, nr, nc, "u8", 2, 16)
    o.vo_fill_border_mirror(h.ptr())
    gp = "vfloat2" if as_float else "vint2"
    g1, g2 = orc.HostImage(nr, nc, gp), orc.HostImage(nr, nc, gp)
    ref.vppref_scharr_u8(h.ptr(), g1.ptr(), int(as_float)); o.vo_scharr_u8(h.ptr(), g2.ptr(), int(as_float))
    assert np.array_equal(g1.get().view(np.int32), g2.get().view(np.int32))
    l1, l2 = orc.HostImage(nr, nc, "u8"), orc.HostImage(nr, nc, "u8")
    ref.vppref_lowpass_u8(h.ptr(), l1.ptr()); o.vo_lowpass(h.ptr(), l2.ptr(), 0)
    assert np.array_equal(l1.get(), l2.get())


@settings(**SET)
@given(kr=st.integers(4, 12), kc=st.integers(4, 12), seed=st.integers(0, 2 ** 16), kind=st.sampled_from([0, 1, 2]), border=st.integers(2, 4))
def test_pyramids_odd_safe_sizes(libs, kr, kc, seed, kind, border):
    """sizes 4k+1 keep all three levels free of the reference's uninitialised low-pass border"""
    ref, o = libs
    from tests.oracle_ops import oracle_grad_pyramid, oracle_pyramid

    nr, nc = 4 * kr + 1, 4 * kc + 1
    a = np.random.default_rng(seed).integers(0, 256, (nr, nc), dtype=np.uint8)
    src = orc.HostImage(nr, nc, "u8", data=a)
    pix = ["u8", "vint2", "vfloat2"][kind]
    mine = oracle_pyramid(a, 3, "u8", border, o)
    if kind:
        mine = oracle_grad_pyramid(mine, pix, border, o)
    theirs = [orc.HostImage(l.nrows, l.ncols, pix, border=border) for l in mine]
    ref.vppref_pyramid(src.ptr(), 3, orc.desc_array(theirs), kind)
    for lvl in range(3):
        x, y = theirs[lvl].get(True), mine[lvl].get(True)
        if pix != "u8":
            x, y = x.view(np.int32), y.view(np.int32)
        assert np.array_equal(x, y), lvl


@settings(**SET)
@given(nr=st.integers(8, 48), nc=st.integers(8, 70), th=st.integers(0, 80), seed=st.integers(0, 2 ** 16), mode=st.sampled_from([0, 1, 2]),
       bs=st.integers(2, 12), maskval=st.sampled_from([None, 0xFF, 0x01, 0x10, 0x11, 0x80]), levels=st.sampled_from([2, 3, 256]))
def test_fast9_any_geometry(libs, nr, nc, th, seed, mode, bs, maskval, levels):
    ref, o = libs
    r = np.random.default_rng(seed)
    img = (r.integers(0, levels, (nr, nc)) * (255 // max(levels - 1, 1))).astype(np.uint8)
    h = orc.HostImage(nr, nc, "u8", border=3, aligned=32, data=img, fill_border="mirror")
    hm = None
    if maskval is not None:
        m = (r.integers(0, 2, (nr, nc)) * maskval).astype(np.uint8)
        hm = orc.HostImage(nr, nc, "u8", aligned=32, data=m)
    k1, k2 = np.zeros((img.size, 2), np.int32), np.zeros((img.size, 2), np.int32)
    s1, s2 = np.zeros(img.size, np.int32), np.zeros(img.size, np.int32)
    n1 = ref.vppref_fast9_u8(h.ptr(), th, hm.ptr() if hm else None, mode, bs, k1.ctypes.data, s1.ctypes.data, img.size)
    n2 = o.vo_fast9_u8(h.ptr(), th, hm.ptr() if hm else None, mode, bs, 0, k2.ctypes.data, s2.ctypes.data, img.size)
    assert n1 == n2
    if mode == 2:
        order = np.lexsort((k2[:n2, 1], k2[:n2, 0]))
        k2[:n2], s2[:n2] = k2[:n2][order], s2[:n2][order]
    assert np.array_equal(k1[:n1], k2[:n2]) and np.array_equal(s1[:n1], s2[:n2])


@settings(max_examples=300, deadline=None, derandomize=True, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow])
@given(seed=st.integers(0, 2 ** 16), winsize=st.sampled_from([5, 7, 9, 11, 15]), nscales=st.sampled_from([1, 2]), niter=st.integers(1, 30),
       sr=st.floats(-3, 3), sc=st.floats(-3, 3), min_ev=st.sampled_from([0.0001, 0.01, 1.0]), delta=st.sampled_from([0.01, 0.1, 0.5]))
def test_lucas_kanade_any_parameters(libs, seed, winsize, nscales, niter, sr, sc, min_ev, delta):
    """reference lucas_kanade() vs the oracle, bit for bit, over random scenes, window sizes, iteration caps and thresholds.
    winsize 3 is left out: lucas_kanade.hpp:149 gives the pyramids a border of winsize/2 = 1, and the 5-tap low-pass of
    pyramid.hh:179-181 then reads 2 pixels out — past the border, undefined values (the CUDA path refuses it: VPPB_E_BORDER)."""
    ref, o = libs
    from tests import scenes
    from tests.oracle_ops import oracle_lucas_kanade

    nr, nc = 101 + 2 * (seed % 20), 121 + 2 * (seed % 17)
    f1, f2, pts = scenes.lk_pair(nr, nc, 60, seed=seed, shift=(sr, sc), margin=30)
    h1, h2 = orc.HostImage(nr, nc, "u8", data=f1), orc.HostImage(nr, nc, "u8", data=f2)
    n = len(pts)
    flow, dist = np.zeros((n, 2), np.float32), np.zeros(n, np.float32)
    ref.vppref_lucas_kanade(h1.ptr(), h2.ptr(), pts.ctypes.data, None, n, niter, winsize, nscales, min_ev, delta, flow.ctypes.data, dist.ctypes.data)
    rflow, rdist = oracle_lucas_kanade(f1, f2, pts, niterations=niter, winsize=winsize, nscales=nscales, min_ev=min_ev, delta=delta, lib=o)
    # A track that diverges wanders to the image edge, where the reference's un-checked bilinear taps read past the
    # allocated border (undefined values; the oracle clamps) - it may even come back.  Such tracks are recognisable by
    # their size: the synthetic motion is <= 3 px.  Every track that stayed small on both sides and ends well inside the
    # image must match bit for bit; the others must be rare.
    end = pts + rflow
    m = winsize // 2 + 2
    inside = (end[:, 0] >= m) & (end[:, 0] <= nr - 1 - m) & (end[:, 1] >= m) & (end[:, 1] <= nc - 1 - m) & np.isfinite(end).all(axis=1)
    small = (np.abs(np.nan_to_num(flow, nan=1e9, posinf=1e9, neginf=1e9)).max(axis=1) <= 8) & (np.abs(np.nan_to_num(rflow, nan=1e9, posinf=1e9, neginf=1e9)).max(axis=1) <= 8)
    same = (flow.view(np.int32) == rflow.view(np.int32)).all(axis=1) & (dist.view(np.int32) == rdist.view(np.int32))
    sane = inside & small
    assert sane.mean() > 0.5
    assert same[sane].all(), np.abs(flow - rflow)[sane & ~same].max()
    assert (~same).mean() < 0.1


@settings(max_examples=150, deadline=None, derandomize=True, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow])
@given(seed=st.integers(0, 2 ** 16), ws=st.sampled_from([5, 7, 9, 11]), nscales=st.integers(1, 3), min_scale=st.integers(0, 1), prop=st.integers(0, 3),
       patch=st.sampled_from([3, 5]), nk=st.integers(1, 400))
def test_semi_dense_flow_any_parameters(libs, seed, ws, nscales, min_scale, prop, patch, nk):
    """random keypoint sets (duplicates and several keypoints per cell included, any order): the first-claim,
    Gauss-Seidel propagation and reporting rules of the reference's serial build, bit for bit"""
    ref, o = libs
    from tests import scenes

    if min_scale >= nscales:
        min_scale = nscales - 1
    nr, nc = 97, 129  # 2^5 * m + 1: every level odd
    r = np.random.default_rng(seed)
    f1, f2, _ = scenes.lk_pair(nr, nc, 4, seed=seed, shift=(float(r.integers(-3, 4)), float(r.integers(-3, 4))), margin=10)
    kps = np.stack([r.integers(0, nr, nk), r.integers(0, nc, nk)], axis=1).astype(np.int32)
    h1, h2 = orc.HostImage(nr, nc, "u8", data=f1), orc.HostImage(nr, nc, "u8", data=f2)
    res = []
    for fn in (ref.vppref_semi_dense_flow, o.vo_semi_dense_flow):
        pos, dist, valid = np.zeros((nk, 2), np.int32), np.zeros(nk, np.int32), np.zeros(nk, np.uint8)
        fn(h1.ptr(), h2.ptr(), kps.ctypes.data, nk, ws, nscales, min_scale, prop, patch, pos.ctypes.data, dist.ctypes.data, valid.ctypes.data)
        res.append((pos, dist, valid))
    assert np.array_equal(res[0][2], res[1][2])
    ok = res[0][2] > 0
    assert np.array_equal(res[0][0][ok], res[1][0][ok])
    assert np.array_equal(res[0][1][ok], res[1][1][ok])
